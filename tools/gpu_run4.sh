timeout 400 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 4 -c 1 -f -o gpurun_out/r01_filter_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>gpurun_out/ncu_filter.err
ls -la gpurun_out/*.ncu-rep
