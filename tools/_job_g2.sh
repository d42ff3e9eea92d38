mkdir -p gpurun_out
G=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/tests_dist.log 2>&1; echo "tests_dist rc=$?"; grep -v "timed out" gpurun_out/tests_dist.log | tail -3 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_weak_g$G.json 2> gpurun_out/bench_weak_g$G.err; echo "bench weak G=$G rc=$?"; head -c 400 gpurun_out/bench_weak_g$G.json; grep -v "timed out" gpurun_out/bench_weak_g$G.err | tail -3 | cut -c1-300
