// K10 — CTC prefix beam search with bigram LM, one thread block per utterance.
//
// Replaces the pure-Python search of timit/utils/BeamSearch.py:73-153 (called from
// BeamDecoder.decode, timit/utils/ctcDecoder.py:181-192) and follows its arithmetic exactly so that the
// decoded label sequences are identical:
//   * scores are float64 built from log() of the *float32* probabilities, added in the reference's order
//     (log p + lm*alpha) + base; log_add(x,y) = max + log(1 + exp(min - max)) with the absorbing
//     sentinel LOG_ZERO = -99999999.0;
//   * a frame is skipped when (1 - p_blank) < 0.1 and a repeated label continues from the blank-ending
//     score when p_{t-1}(blank) < 0.9, both compared in float32;
//   * the next beam is the first `beam_width` entries of a *stable* descending sort of all candidates of
//     the previous frame, ties resolved by dictionary insertion order = candidate position
//     (beam rank major; the "stay" candidate first, then extensions by class index);
//   * a prefix reached both by staying (y) and by extending its parent (y[:-1] + k) is one entry whose
//     scores are log-added (each entry has at most these two contributions, and log_add is commutative
//     bit-for-bit, so the merge order cannot change the result);
//   * end of utterance: add alpha * bigram(last, </s>), divide the log score by the label count, arg-max.
// Reference failure modes are reported through `status` so the Python shim can raise the same exception:
// 1 = IndexError (empty prefix in the final beam), 2 = ValueError (log of a zero probability),
// 3 = KeyError (unit pair missing from the LM).
//
// Device mapping: candidates (beam x class) are spread over the 1024 threads of the block, the candidate
// keys live in shared memory; the best W are picked by a radix select + a small bitonic sort (select_top), beam records (scores, last label,
// length, 64-bit prefix hash) are double-buffered in shared memory, prefix equality for the merge is
// found through a shared-memory hash table and then verified exactly on the stored label sequences.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

constexpr int BEAM_THREADS = 1024;
constexpr double LOG_ZERO = -99999999.0;

__device__ __forceinline__ double log_add(double x, double y) {
    if (x <= LOG_ZERO) return y;
    if (y <= LOG_ZERO) return x;
    if (y - x > 0.0) { double t = x; x = y; y = t; }
    return x + log(1.0 + exp(y - x));
}

__device__ __forceinline__ unsigned long long mix_hash(unsigned long long h, int k) {
    unsigned long long z = h ^ (static_cast<unsigned long long>(k) + 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z ? z : 1ULL;
}

// a sorts before b: larger key first, equal keys by smaller position (stable order of the reference)
__device__ __forceinline__ bool before(double ka, int pa, double kb, int pb) { return ka > kb || (ka == kb && pa < pb); }

struct BeamRec {
    double* total; double* nonblank; double* blank;
    int* last; int* len;
    unsigned long long* hash; unsigned long long* phash;
};

struct BeamParams {
    const float* probs;       // [N, T, C] float32 probabilities
    const int64_t* lengths;   // [N]
    const double* lm;         // [(C+1), (C+1)]
    double lm_alpha;
    int T, N, C, W, blank, P2;  // P2 = candidate capacity >= W*C
    int lm_in_smem;           // 1: the kernel copies the bigram table into shared memory
    int* seq;                 // [N, 2, W, T]
    int* out_labels;          // [N, T]
    int* out_len;             // [N]
    int* status;              // [N]
    long long* trace;         // debug (CTCB200_BEAM_TRACE=1): per-phase cycle totals of block 0, or null
};

__device__ void bitonic_sort(double* key, int* pos, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double ka = key[i], kb = key[ixj];
                    const int pa = pos[i], pb = pos[ixj];
                    const bool up = (i & k) == 0;  // this run sorts "best first"
                    const bool swap = up ? before(kb, pb, ka, pa) : before(ka, pa, kb, pb);
                    if (swap) { key[i] = kb; key[ixj] = ka; pos[i] = pb; pos[ixj] = pa; }
                }
            }
            __syncthreads();
        }
    }
}

// Order-preserving map double -> uint64 (larger value = larger integer); -0.0 counts as +0.0 like the == of the reference's sort
__device__ __forceinline__ unsigned long long key_bits(double k) {
    if (k == 0.0) k = 0.0;
    const long long b = __double_as_longlong(k);
    const unsigned long long u = static_cast<unsigned long long>(b);
    return (b < 0) ? ~u : (u | 0x8000000000000000ULL);
}

constexpr int TOP_MAX = 256;   // beam widths up to 256 (BeamDecoder's default is 200)

// The next beam is the first W entries of a stable descending sort of ALL candidates (BeamSearch.py:29-33,96): only those W
// are needed, so instead of sorting the whole pool (8192 keys: 91 compare-exchange stages, 88 % of the kernel's time in
// round 1's profile) the W-th largest key is found by an 8-pass radix select over the 64-bit key images, the entries above
// it plus the first (lowest-position) ties at it are compacted into a 256-entry buffer, and only that buffer is sorted by
// (key descending, position ascending). Candidates sit at their own position in skey before the selection (spos[i] == i), so
// "position" is the array index. Result: skey[0 .. want), spos[0 .. want) exactly as the full sort would have left them.
__device__ void select_top(double* skey, int* spos, int n, int want, double* okey, int* opos, int* hist, int* sscan) {
    // hist: [8 passes][256 bins], zeroed here once; sscan: 2 + 32 + 2 ints
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    if (want > n) want = n;
    if (want <= 0) return;
    const int ept = (n + nthreads - 1) / nthreads;
    const int lo = tid * ept, hi = min(n, lo + ept);
    for (int i = tid; i < 8 * 256; i += nthreads) hist[i] = 0;
    if (tid == 0) { sscan[0] = 0; sscan[1] = 0; }
    // key images of this thread's candidates stay in registers for all passes (the pool holds at most KEYS_PER_THREAD per
    // thread: P2 <= 8192 slots over 1024 threads)
    constexpr int KPT = 8;
    unsigned long long u[KPT];
    unsigned long long diff = 0ULL;
    const unsigned long long u_first = key_bits(skey[0]);
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const int i = lo + j;
        u[j] = (j < ept && i < hi) ? key_bits(skey[i]) : u_first;
        diff |= u[j] ^ u_first;
    }
    for (int i = lo + KPT; i < hi; ++i) diff |= key_bits(skey[i]) ^ u_first;   // wider pools: the tail is re-read every pass
    __syncthreads();   // hist / sscan initialised
    // bits in which the candidates differ at all: a byte position where they all agree needs no counting pass
    diff |= __shfl_xor_sync(0xffffffffu, diff, 16);
    diff |= __shfl_xor_sync(0xffffffffu, diff, 8);
    diff |= __shfl_xor_sync(0xffffffffu, diff, 4);
    diff |= __shfl_xor_sync(0xffffffffu, diff, 2);
    diff |= __shfl_xor_sync(0xffffffffu, diff, 1);
    if (lane == 0 && diff != 0ULL) {
        atomicOr(reinterpret_cast<unsigned int*>(&sscan[0]), static_cast<unsigned int>(diff));
        atomicOr(reinterpret_cast<unsigned int*>(&sscan[1]), static_cast<unsigned int>(diff >> 32));
    }
    __syncthreads();
    diff = static_cast<unsigned long long>(static_cast<unsigned int>(sscan[0])) |
           (static_cast<unsigned long long>(static_cast<unsigned int>(sscan[1])) << 32);
    unsigned long long prefix = 0ULL, done_mask = 0ULL;   // done_mask: the key bits decided so far
    int remaining = want;
    bool exact_cut = false;   // the want-th and (want+1)-th key already differ in the decided bits: no further pass needed
#pragma unroll 1
    for (int pass = 7; pass >= 0 && !exact_cut; --pass) {
        const int shift = pass * 8;
        const unsigned long long mask_above = done_mask;
        done_mask |= 255ULL << shift;
        if (((diff >> shift) & 255ULL) == 0ULL) {      // every candidate has the same byte here
            prefix |= u_first & (255ULL << shift);
            continue;
        }
        int* h = hist + pass * 256;
        auto count = [&](int i, unsigned long long ui) {
            int d = -1;
            if (i < hi && (ui & mask_above) == prefix) d = static_cast<int>((ui >> shift) & 255ULL);
            // the high bytes of scores of similar magnitude coincide across the whole warp: one atomic for 32 lanes then;
            // otherwise plain shared-memory atomics (few lanes still match the prefix in the later passes)
            const int d0 = __shfl_sync(0xffffffffu, d, 0);
            if (__all_sync(0xffffffffu, d == d0)) {
                if (lane == 0 && d0 >= 0) atomicAdd(&h[d0], 32);
            } else if (d >= 0) {
                atomicAdd(&h[d], 1);
            }
        };
        // uniform trip counts: the warp-wide vote needs every lane present
#pragma unroll
        for (int j = 0; j < KPT; ++j)
            if (j < ept) count(lo + j, u[j]);
        for (int j = KPT; j < ept; ++j) count(lo + j, (lo + j < hi) ? key_bits(skey[lo + j]) : 0ULL);
        __syncthreads();
        // EVERY warp finds the digit itself (256 bins, 8 per lane, one shuffle scan): no second barrier for a broadcast
        int loc[8], sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { loc[b] = h[lane * 8 + b]; sum += loc[b]; }
        int x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_down_sync(0xffffffffu, x, o);
            if (lane + o < 32) x += y;
        }
        const int above = x - sum;   // entries whose digit lies in a higher lane's bins
        int digit = -1, gt = 0, cnt = 0;
        if (above < remaining && remaining <= above + sum) {
            int run = above;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                if (run < remaining && remaining <= run + loc[b]) { digit = lane * 8 + b; gt = run; cnt = loc[b]; }
                run += loc[b];
            }
        }
        const unsigned owner = __ballot_sync(0xffffffffu, digit >= 0);   // exactly one lane holds the bin
        const int src = __ffs(owner) - 1;
        digit = __shfl_sync(0xffffffffu, digit, src);
        gt = __shfl_sync(0xffffffffu, gt, src);
        cnt = __shfl_sync(0xffffffffu, cnt, src);
        prefix |= static_cast<unsigned long long>(digit) << shift;
        remaining -= gt;
        // the whole bin is taken: the selected set is exactly "decided bits >= prefix", whatever the lower bits are
        exact_cut = (cnt == remaining);
    }
    // the want-th largest key agrees with `prefix` in the decided bits; `remaining` entries equal to it there are taken, lowest
    // positions first (after a full run the decided bits are all 64 and this is the tie rule of the stable sort; after an early
    // exit every such entry is taken, so their order does not matter)
    const unsigned long long kth = prefix;
    int cg = 0, ce = 0;
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        if (j < ept && lo + j < hi) {
            const unsigned long long um = u[j] & done_mask;
            cg += um > kth;
            ce += um == kth;
        }
    }
    for (int i = lo + KPT; i < hi; ++i) {
        const unsigned long long um = key_bits(skey[i]) & done_mask;
        cg += um > kth;
        ce += um == kth;
    }
    // block-wide exclusive scan of (cg, ce) packed into one int (both < 2^15)
    int packed = cg | (ce << 16), inc = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sscan[2 + warp] = inc;
    __syncthreads();
    int wtot = lane < (nthreads >> 5) ? sscan[2 + lane] : 0, wsum = wtot;    // every warp scans the 32 warp totals itself
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, wsum, o);
        if (lane >= o) wsum += y;
    }
    const int warp_excl = __shfl_sync(0xffffffffu, wsum - wtot, warp);
    const int excl = inc - packed + warp_excl;
    int bg = excl & 0xffff, be = excl >> 16;
    const int n_gt = want - remaining;
    auto place = [&](int i, unsigned long long ui) {
        const unsigned long long um = ui & done_mask;
        if (um > kth) { okey[bg] = skey[i]; opos[bg] = i; ++bg; }
        else if (um == kth) {
            if (be < remaining) { okey[n_gt + be] = skey[i]; opos[n_gt + be] = i; }
            ++be;
        }
    };
#pragma unroll
    for (int j = 0; j < KPT; ++j)
        if (j < ept && lo + j < hi) place(lo + j, u[j]);
    for (int i = lo + KPT; i < hi; ++i) place(i, key_bits(skey[i]));
    int wp = 1;
    while (wp < want) wp <<= 1;
    for (int i = want + tid; i < wp; i += nthreads) { okey[i] = -INFINITY; opos[i] = 0x7fffffff; }
    __syncthreads();
    // sort the <= 256 selected entries by (key descending, position ascending): one thread per pair, the first wp/2 threads
    // only (<= 4 warps), synchronised with a named barrier instead of the whole block
    const int half = wp >> 1;
    const int team = half > 32 ? half : 32;     // whole warps take part in the barriers; threads >= half carry no pair
    if (tid < team) {
        for (int k = 2; k <= wp; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (tid < half) {
                    const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));   // index with bit j clear
                    const int ixj = i | j;
                    const double ka = okey[i], kb = okey[ixj];
                    const int pa = opos[i], pb = opos[ixj];
                    const bool up = (i & k) == 0;
                    const bool swap = up ? before(kb, pb, ka, pa) : before(ka, pa, kb, pb);
                    if (swap) { okey[i] = kb; okey[ixj] = ka; opos[i] = pb; opos[ixj] = pa; }
                }
                if (team > 32) named_bar_sync(1, team);
                else __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < wp; i += nthreads) { skey[i] = okey[i]; spos[i] = opos[i]; }
    __syncthreads();
}

__global__ void __launch_bounds__(BEAM_THREADS, 1) beam_search_kernel(BeamParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int C = p.C, W = p.W, T = p.T, blank = p.blank, P2 = p.P2;
    // ---- shared memory carve-up ----
    double* skey = reinterpret_cast<double*>(smem_raw);
    double* rec_d = skey + P2;                       // 2 buffers x 3 arrays x W doubles
    double* logp = rec_d + 6 * W;                    // C
    unsigned long long* rec_h = reinterpret_cast<unsigned long long*>(logp + C);  // 2 x 2 x W
    int HT = 1;
    while (HT < 2 * W) HT <<= 1;
    unsigned long long* ht_key = rec_h + 4 * W;      // HT
    int* spos = reinterpret_cast<int*>(ht_key + HT); // P2
    int* rec_i = spos + P2;                          // 2 x 2 x W
    int* ht_val = rec_i + 4 * W;                     // HT
    int* merged_with = ht_val + HT;                  // W
    float* prow = reinterpret_cast<float*>(merged_with + W);  // C
    // What a pool entry stands for is DERIVED from its position unless it is the "stay" entry of a beam: tag[pos] = r >= 0 marks
    // the entry that carries beam r unchanged (its own slot r*C, or the position of the extension it merged with), with
    // stay_nb / stay_bl holding that entry's non-blank / blank scores; tag = -1: extension (parent pos / C, label from the slot,
    // non-blank score = the key itself, blank score = log 0). Nothing about the pool lives in global memory.
    double* stay_nb = reinterpret_cast<double*>(reinterpret_cast<uintptr_t>(prow + C + 1) & ~uintptr_t(7)) + 1;   // W (8-byte aligned)
    double* stay_bl = stay_nb + W;                   // W
    double* lm_s = stay_bl + W;                      // (C+1)^2 when p.lm_in_smem, else unused
    short* tag = reinterpret_cast<short*>(lm_s + (p.lm_in_smem ? (C + 1) * (C + 1) : 0));   // P2
    __shared__ int s_flags[4];                       // [0] nbeams, [1] status, [2] processed frames, [3] skip
    __shared__ double s_okey[TOP_MAX];               // selection buffer of select_top
    __shared__ int s_opos[TOP_MAX], s_hist[8 * 256], s_scan[40];

    auto rec = [&](int b) {
        BeamRec r;
        r.total = rec_d + b * 3 * W; r.nonblank = r.total + W; r.blank = r.nonblank + W;
        r.last = rec_i + b * 2 * W; r.len = r.last + W;
        r.hash = rec_h + b * 2 * W; r.phash = r.hash + W;
        return r;
    };

    const float* probs_n = p.probs + static_cast<size_t>(n) * T * C;
    const double* lm = p.lm;
    if (p.lm_in_smem) {   // the bigram table (31.7 KB at C = 62) is read once per candidate: keep it on chip
        for (int i = tid; i < (C + 1) * (C + 1); i += blockDim.x) lm_s[i] = p.lm[i];
        lm = lm_s;
    }
    int* seq_base = p.seq + static_cast<size_t>(n) * 2 * W * T;
    int len_n = static_cast<int>(p.lengths[n]);
    if (len_n > T) len_n = T;

    // root beam: the empty prefix with prBlank = prTotal = 0
    int cur = 0;  // index of the beam-record buffer that holds the current beams
    if (tid == 0) {
        BeamRec r = rec(0);
        r.total[0] = 0.0; r.nonblank[0] = LOG_ZERO; r.blank[0] = 0.0;
        r.last[0] = -1; r.len[0] = 0; r.hash[0] = 0x243F6A8885A308D3ULL; r.phash[0] = 0ULL;
        s_flags[0] = 1; s_flags[1] = 0; s_flags[2] = 0;
    }
    __syncthreads();
    bool have_pool = false;  // candidates of an earlier frame are waiting to be ranked
    int pool_n = 0;
    // phase timing (block 0, thread 0): [0] sort, [1] beam records + sequence copies, [2] frame setup + hash table,
    // [3] extensions, [4] stays / merge, [5] frames processed
    long long tr[6] = {0, 0, 0, 0, 0, 0};
    const bool tracing = p.trace != nullptr && n == 0 && tid == 0;
#define BTICK(var) long long var = tracing ? clock64() : 0

    // Rank the waiting candidates and turn the best W into the current beam records (+ their label sequences).
    auto select_beams = [&]() {
        __syncthreads();
        BTICK(t_s0);
        select_top(skey, spos, pool_n, W, s_okey, s_opos, s_hist, s_scan);
        BTICK(t_s1);
        tr[0] += t_s1 - t_s0;
        int live = 0;
        // number of live entries among the first W (dead ones carry -inf and sort last)
        if (tid == 0) {
            int c = 0;
            const int lim = W < pool_n ? W : pool_n;
            while (c < lim && skey[c] != -INFINITY) ++c;
            s_flags[0] = c;
        }
        __syncthreads();
        live = s_flags[0];
        const BeamRec prev = rec(cur), nxt = rec(cur ^ 1);
        const int frames_done = s_flags[2];
        const int* seq_old = seq_base + static_cast<size_t>(frames_done & 1) * W * T;
        int* seq_new = seq_base + static_cast<size_t>((frames_done + 1) & 1) * W * T;
        for (int r = tid; r < live; r += blockDim.x) {
            const int pos = spos[r];
            const int tg = tag[pos];
            const int slot = pos % C;
            const int parent = tg >= 0 ? tg : pos / C;
            const int label = tg >= 0 ? -1 : ((slot - 1 < blank) ? slot - 1 : slot);
            nxt.total[r] = skey[r];
            nxt.nonblank[r] = tg >= 0 ? stay_nb[tg] : skey[r];
            nxt.blank[r] = tg >= 0 ? stay_bl[tg] : LOG_ZERO;
            nxt.len[r] = prev.len[parent] + (label >= 0 ? 1 : 0);
            nxt.last[r] = label >= 0 ? label : prev.last[parent];
            nxt.hash[r] = label >= 0 ? mix_hash(prev.hash[parent], label) : prev.hash[parent];
            nxt.phash[r] = label >= 0 ? prev.hash[parent] : prev.phash[parent];
        }
        // label sequences: one warp per selected beam copies its parent's labels and appends
        const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
        for (int r = warp; r < live; r += nwarps) {
            const int pos = spos[r];
            const int tg = tag[pos];
            const int slot = pos % C;
            const int parent = tg >= 0 ? tg : pos / C;
            const int label = tg >= 0 ? -1 : ((slot - 1 < blank) ? slot - 1 : slot);
            const int plen = prev.len[parent];
            for (int e = lane; e < plen; e += 32) seq_new[static_cast<size_t>(r) * T + e] = seq_old[static_cast<size_t>(parent) * T + e];
            if (lane == 0 && label >= 0) seq_new[static_cast<size_t>(r) * T + plen] = label;
        }
        __syncthreads();
        cur ^= 1;
        if (tid == 0) s_flags[2] = frames_done + 1;
        __syncthreads();
        BTICK(t_s2);
        tr[1] += t_s2 - t_s1;
    };

    for (int t = 0; t < len_n; ++t) {
        for (int c = tid; c < C; c += blockDim.x) prow[c] = probs_n[static_cast<size_t>(t) * C + c];
        __syncthreads();
        if ((1.0f - prow[blank]) < 0.1f) { __syncthreads(); continue; }  // near-certain blank: frame skipped
        if (have_pool) select_beams();
        BTICK(t_f0);
        const int nbeams = s_flags[0];
        const BeamRec b = rec(cur);
        const int* seq_cur = seq_base + static_cast<size_t>(s_flags[2] & 1) * W * T;
        for (int c = tid; c < C; c += blockDim.x) {
            const float pv = prow[c];
            if (pv == 0.0f) s_flags[1] = 2;  // math.log(0.0) -> ValueError in the reference
            logp[c] = log(static_cast<double>(pv));
        }
        for (int i = tid; i < HT; i += blockDim.x) { ht_key[i] = 0ULL; ht_val[i] = -1; }
        for (int i = tid; i < nbeams; i += blockDim.x) merged_with[i] = -1;
        __syncthreads();
        for (int r = tid; r < nbeams; r += blockDim.x) {
            const unsigned long long h = b.hash[r];
            unsigned int slot = static_cast<unsigned int>(h) & (HT - 1);
            while (true) {
                const unsigned long long old = atomicCAS(&ht_key[slot], 0ULL, h);
                if (old == 0ULL) { ht_val[slot] = r; break; }
                slot = (slot + 1) & (HT - 1);  // distinct prefixes with equal hash simply occupy two slots
            }
        }
        __syncthreads();
        const int tprev = (t - 1 + T) % T;
        const bool prev_blank_lt = probs_n[static_cast<size_t>(tprev) * C + blank] < 0.9f;
        pool_n = nbeams * C;
        BTICK(t_f1);
        tr[2] += t_f1 - t_f0;
        // pass 1: extensions
        for (int pos = tid; pos < pool_n; pos += blockDim.x) {
            const int i = pos / C, slot = pos - i * C;
            if (slot == 0) continue;
            const int k = (slot - 1 < blank) ? slot - 1 : slot;
            const int last_i = b.last[i], len_i = b.len[i];
            const double lmv = lm[static_cast<size_t>(len_i ? last_i : C) * (C + 1) + k];
            if (lmv != lmv) s_flags[1] = 3;  // unit pair unknown to the LM -> KeyError in the reference
            const double lm_term = lmv * p.lm_alpha;
            const double base = (len_i && last_i == k && prev_blank_lt) ? b.blank[i] : b.total[i];
            const double score = logp[k] + lm_term + base;
            skey[pos] = score;
            tag[pos] = -1;
            // does y_i + (k,) coincide with a beam prefix y_j (which contributes its own "stay" entry)?
            const unsigned long long h = mix_hash(b.hash[i], k);
            unsigned int hs = static_cast<unsigned int>(h) & (HT - 1);
            while (ht_key[hs] != 0ULL) {
                if (ht_key[hs] == h) {
                    const int j = ht_val[hs];
                    if (j >= 0 && b.len[j] == len_i + 1 && b.last[j] == k && b.phash[j] == b.hash[i]) {
                        bool same = true;
                        for (int e = 0; e < len_i; ++e)
                            if (seq_cur[static_cast<size_t>(j) * T + e] != seq_cur[static_cast<size_t>(i) * T + e]) { same = false; break; }
                        if (same) { merged_with[j] = pos; break; }
                    }
                }
                hs = (hs + 1) & (HT - 1);
            }
        }
        __syncthreads();
        BTICK(t_f2);
        tr[3] += t_f2 - t_f1;
        // pass 2: stays (and the merge with a coinciding extension)
        for (int r = tid; r < nbeams; r += blockDim.x) {
            double nb = LOG_ZERO;
            if (b.len[r]) nb = b.nonblank[r] + logp[b.last[r]];
            const double bl = b.total[r] + logp[blank];
            double tot = log_add(bl, nb);
            const int ps = r * C, pe = merged_with[r];
            int home = ps;
            if (pe >= 0) {
                const double score = skey[pe];
                nb = log_add(nb, score);
                tot = log_add(tot, score);
                home = pe < ps ? pe : ps;                 // the entry keeps its first insertion position
                const int dead = pe < ps ? ps : pe;
                skey[dead] = -INFINITY;
                tag[dead] = -1;
            }
            skey[home] = tot;
            tag[home] = static_cast<short>(r);
            stay_nb[r] = nb;
            stay_bl[r] = bl;
        }
        have_pool = true;
        __syncthreads();
        BTICK(t_f3);
        tr[4] += t_f3 - t_f2;
        tr[5] += 1;
        if (s_flags[1]) break;
    }
    if (tracing)
        for (int i = 0; i < 6; ++i) p.trace[i] = tr[i];

    int status = s_flags[1];
    if (status == 0) {
        if (have_pool) select_beams();
        const int nbeams = s_flags[0];
        const BeamRec b = rec(cur);
        // final LM step over the best W prefixes, length normalisation, arg-max (first best wins)
        for (int r = tid; r < nbeams; r += blockDim.x) {
            if (b.len[r] == 0) { s_flags[1] = 1; skey[r] = -INFINITY; spos[r] = r; continue; }  // classes[y[-1]] on ()
            const double lmv = lm[static_cast<size_t>(b.last[r]) * (C + 1) + C];
            if (lmv != lmv) s_flags[1] = 3;
            const double eos = b.total[r] + lmv * p.lm_alpha;
            skey[r] = eos * (1.0 / static_cast<double>(b.len[r]));
            spos[r] = r;
        }
        __syncthreads();
        status = s_flags[1];
        if (status == 0) {
            if (tid == 0) {
                int best = 0;
                for (int r = 1; r < nbeams; ++r)
                    if (skey[r] > skey[best]) best = r;
                s_flags[3] = best;
            }
            __syncthreads();
            const int best = s_flags[3];
            const int* seq_cur = seq_base + static_cast<size_t>(s_flags[2] & 1) * W * T;
            const int blen = b.len[best];
            for (int e = tid; e < blen; e += blockDim.x) p.out_labels[static_cast<size_t>(n) * T + e] = seq_cur[static_cast<size_t>(best) * T + e];
            if (tid == 0) p.out_len[n] = blen;
        }
    }
    if (tid == 0) {
        p.status[n] = status;
        if (status) p.out_len[n] = 0;
    }
}

// out[n, t, c] = expf(in[t, n, c]) — the torch.exp(probs.transpose(0,1)) of ctcDecoder.py:189-190
__global__ void exp_transpose_kernel(const float* __restrict__ lp, float* __restrict__ out, int T, int N, int C) {
    const long long total = static_cast<long long>(T) * N * C;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long tn = e / C;
        const int nn = static_cast<int>(tn % N), t = static_cast<int>(tn / N);
        out[(static_cast<long long>(nn) * T + t) * C + c] = expf(lp[e]);
    }
}

size_t beam_smem_bytes(int W, int C, int P2, bool lm_in_smem) {
    int HT = 1;
    while (HT < 2 * W) HT <<= 1;
    size_t b = 0;
    b += sizeof(double) * (static_cast<size_t>(P2) + 6 * W + C);
    b += sizeof(unsigned long long) * (4 * static_cast<size_t>(W) + HT);
    b += sizeof(int) * (static_cast<size_t>(P2) + 4 * W + HT + W);
    b += sizeof(float) * (C + 1) + 16;                                   // prow + alignment of what follows
    b += sizeof(double) * (2 * static_cast<size_t>(W) + (lm_in_smem ? static_cast<size_t>(C + 1) * (C + 1) : 0));   // stay_nb, stay_bl, LM
    b += sizeof(short) * static_cast<size_t>(P2);                        // tag
    return b + 16;   // + static shared memory of the kernel (selection buffers, ~4.2 KB)
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int64_t ctcb200_beam_workspace_bytes(int T, int N, int C, int beam_width) {
    (void)C;   // the candidate pool lives in shared memory; the workspace holds the two label-sequence buffers of every utterance
    return N * (static_cast<int64_t>(2) * beam_width * T * 4) + 256;
}

extern "C" CTCB200_API int ctcb200_exp_transpose(const float* log_probs_tnc, float* probs_ntc, int T, int N, int C,
                                                 ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0 && C > 0, "exp_transpose: empty shape");
    long long total = static_cast<long long>(T) * N * C;
    long long blocks = (total + 1023) / 1024;
    long long cap = static_cast<long long>(device_sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    exp_transpose_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(log_probs_tnc, probs_ntc, T, N, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_beam_search(const float* probs_ntc, const int64_t* lengths, const double* lm_table,
                                               double lm_alpha, int T, int N, int C, int beam_width, int blank,
                                               void* workspace, int* out_labels, int* out_lengths, int* status,
                                               ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0 && C > 1 && beam_width > 0, "beam_search: bad shape T=%d N=%d C=%d W=%d", T, N, C, beam_width);
    CTCB_REQUIRE(blank >= 0 && blank < C, "beam_search: blank %d out of range", blank);
    CTCB_REQUIRE(beam_width <= TOP_MAX, "beam_search: beam width %d exceeds the supported maximum %d", beam_width, TOP_MAX);
    CTCB_REQUIRE(lm_table != nullptr, "beam_search: a bigram LM table is required (the reference cannot run without one)");
    const int P2 = ((beam_width * C + 31) / 32) * 32;   // candidate capacity (no power-of-two padding: only the top W are sorted)
    // the bigram table goes to shared memory when it fits beside the candidate pool
    int lm_in_smem = 1;
    size_t smem = beam_smem_bytes(beam_width, C, P2, true);
    if (smem > 227 * 1024) { lm_in_smem = 0; smem = beam_smem_bytes(beam_width, C, P2, false); }
    CTCB_REQUIRE(smem <= 227 * 1024, "beam_search: beam_width*classes = %d needs %zu B of shared memory (max 227 KB)",
                 beam_width * C, smem);
    CTCB_CUDA(cudaFuncSetAttribute(beam_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    BeamParams p;
    p.probs = probs_ntc; p.lengths = lengths; p.lm = lm_table; p.lm_alpha = lm_alpha;
    p.T = T; p.N = N; p.C = C; p.W = beam_width; p.blank = blank; p.P2 = P2; p.lm_in_smem = lm_in_smem;
    uint8_t* w = static_cast<uint8_t*>(workspace);
    p.seq = reinterpret_cast<int*>(w);
    p.out_labels = out_labels; p.out_len = out_lengths; p.status = status;
    p.trace = nullptr;
    static long long* dtrace = nullptr;
    if (getenv("CTCB200_BEAM_TRACE")) {   // development aid: phase breakdown of utterance 0 on stderr
        if (!dtrace) CTCB_CUDA(cudaMalloc(&dtrace, 6 * sizeof(long long)));
        CTCB_CUDA(cudaMemsetAsync(dtrace, 0, 6 * sizeof(long long), stream));
        p.trace = dtrace;
    }
    beam_search_kernel<<<N, BEAM_THREADS, smem, stream>>>(p);
    CTCB_LAUNCH_CHECK();
    if (p.trace) {
        long long h[6];
        CTCB_CUDA(cudaStreamSynchronize(stream));
        CTCB_CUDA(cudaMemcpy(h, dtrace, sizeof(h), cudaMemcpyDeviceToHost));
        const double f = h[5] > 0 ? static_cast<double>(h[5]) : 1.0;
        fprintf(stderr, "beam_search trace (utterance 0, %lld unskipped frames, cycles per frame): sort %.0f, beam records + sequence "
                "copies %.0f, frame setup + hash table %.0f, extensions %.0f, stays/merge %.0f\n", h[5], h[0] / f, h[1] / f, h[2] / f,
                h[3] / f, h[4] / f);
    }
    return OK;
}
