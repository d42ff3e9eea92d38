"""Turn gpurun_out/*.csv / *.ncu-rep into the small text summaries committed under profiles/."""
import csv, collections, subprocess, sys, os, io

def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
    agg, tot = collections.OrderedDict(), 0.0
    for d in data:
        k = d["Kernel Name"]
        k = k.replace("ctcb200::<unnamed>::", "").replace("void ", "")[:70]
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e6 if d["Metric Unit"] == "ns" else (v / 1e3 if d["Metric Unit"] == "us" else v)
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v; tot += v
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one cfg2 training step (tools/profile_step.py)\n")
        f.write("# per-launch times under ncu are serialised / cold-cache: compare SHARES, not absolutes\n")
        f.write("total %.3f ms over %d launches\n" % (tot, len(data)))
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s n=%3d %9.3f ms %5.1f%%\n" % (k, c, v, 100 * v / tot))

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "launch__block_size"]

def report(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; selected raw metrics per captured launch (%s)\n" % os.path.basename(rep))
        for r in rows[2:]:
            f.write("\n== %s  grid %s block %s\n" % (r[ix["Kernel Name"]][:100], r[ix.get("Grid Size", 0)], r[ix.get("Block Size", 0)]))
            for w in WANT:
                for h in hdr:
                    if h == w or (h.startswith(w) and h[len(w):len(w) + 1] in ("", ".")):
                        f.write("   %-70s %s %s\n" % (h, r[ix[h]], units[ix[h]]))

def traffic(rep, config, out_json):
    """dram bytes (read + write) per launch of the recurrent kernels -> profiles/ncu_traffic.json (bench.py's roofline.traffic)"""
    import json
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        key = "lstm_bwd_kernel" if "lstm_bwd" in name else ("lstm_fwd_kernel" if "lstm_fwd" in name else None)
        if key is None:
            continue
        rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
        wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
        acc[key].append(rd + wr)
    d = json.load(open(out_json)) if os.path.exists(out_json) else {}
    d.setdefault(config, {})
    for key, v in acc.items():
        d[config][key] = {"dram_bytes": sum(v) / len(v), "launches": len(v),
                          "source": "profiles/%s.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)"
                                    % os.path.basename(rep).replace(".ncu-rep", "")}
    json.dump(d, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    g = "gpurun_out"
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    if os.path.exists("%s/launches_%s.csv" % (g, tag)): launches("%s/launches_%s.csv" % (g, tag), "profiles/launches_%s.txt" % tag)
    for nm in ("prof_lstm_%s" % tag, "prof_gemm_ctc_%s" % tag, "prof_beam_%s" % tag, "prof_misc_%s" % tag):
        if os.path.exists("%s/%s.ncu-rep" % (g, nm)): report("%s/%s.ncu-rep" % (g, nm), "profiles/%s.txt" % nm)
    if os.path.exists("%s/prof_lstm_%s.ncu-rep" % (g, tag)):
        traffic("%s/prof_lstm_%s.ncu-rep" % (g, tag), "cfg2", "profiles/ncu_traffic.json")
