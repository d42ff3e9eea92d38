mkdir -p gpurun_out
timeout 200 python tools/rec_ab_head.py 800 32 512 > gpurun_out/rec_ab_head.json 2> gpurun_out/rec_ab_head.err; echo "rec rc=$?"; cat gpurun_out/rec_ab_head.json; tail -2 gpurun_out/rec_ab_head.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed_input or overlapped_wgrad" > gpurun_out/tests_stream.log 2>&1; echo "tests rc=$?"; grep -v "timed out" gpurun_out/tests_stream.log | tail -8 | cut -c1-400
grep streamed_gx_dg gpurun_out/parity_report.jsonl | cut -c1-600
timeout 300 python tools/gx_stream_ab.py cfg2 bf16 10 > gpurun_out/gx_ab_cfg2.jsonl 2> gpurun_out/gx_ab_cfg2.err; echo "ab rc=$?"; grep -v "timed out" gpurun_out/gx_ab_cfg2.jsonl | cut -c1-200; grep -v "timed out" gpurun_out/gx_ab_cfg2.err | tail -3 | cut -c1-300
