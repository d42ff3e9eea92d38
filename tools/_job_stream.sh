mkdir -p gpurun_out
timeout -s KILL 100 python tools/rec_ab_head.py 800 32 512 > gpurun_out/rec_ab_head.json 2> gpurun_out/rec_ab_head.err; echo "rec rc=$?"; cat gpurun_out/rec_ab_head.json; tail -2 gpurun_out/rec_ab_head.err | cut -c1-300
timeout -s KILL 100 python tools/rec_ab_head.py 1200 64 640 > gpurun_out/rec_ab_head4.json 2> gpurun_out/rec_ab_head4.err; echo "rec4 rc=$?"; cat gpurun_out/rec_ab_head4.json; tail -2 gpurun_out/rec_ab_head4.err | cut -c1-300
