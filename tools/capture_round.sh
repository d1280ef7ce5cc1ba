set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG:-r02_t}_bench_C2.json 2> gpurun_out/${TAG:-r02_t}_bench_C2.err
python bench.py --steps 20 --warmup 3 --config C1 > gpurun_out/${TAG:-r02_t}_bench_C1.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --config C3 > gpurun_out/${TAG:-r02_t}_bench_C3.json 2>/dev/null
python bench.py --steps 30 --warmup 3 --config C5 > gpurun_out/${TAG:-r02_t}_bench_C5.json 2>/dev/null
python bench.py --aux --steps 5 --warmup 3 > gpurun_out/${TAG:-r02_t}_aux.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_fft|k_grain_finish|k_conv2d" -s 80 -c 5 -o gpurun_out/${TAG:-r02_t}_c2 python bench.py --steps 2 --warmup 3 --kernel-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pointwise -s 200 -c 1 -o gpurun_out/${TAG:-r02_t}_c1 python bench.py --steps 2 --warmup 3 --config C1 --kernel-only > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 25 --csv --log-file gpurun_out/${TAG:-r02_t}_launches_c2.csv python bench.py --steps 3 --warmup 3 --kernel-only > /dev/null 2>&1
ls -la gpurun_out | tail -n 12
tail -c 600 gpurun_out/${TAG:-r02_t}_bench_C2.json
