TAG=${1:-fo6}
python tools/gpu_kprobe.py $TAG > gpurun_out/kprobe_$TAG.json 2> gpurun_out/kprobe_$TAG.err; cat gpurun_out/kprobe_$TAG.json; tail -3 gpurun_out/kprobe_$TAG.err
python tools/gpu_lone_rates.py 2>&1 | tail -1
python tools/gpu_wl.py $TAG FHnode,TC 2>&1 | tail -1
