#!/bin/bash
# (the launch list runs with --sm-partition 0: ncu cannot profile kernels launched into green contexts)
# One profiling pass on a GPU box (run through gpurun): the ncu launch list of the bench command, and ncu --set full captures of
# the Hamming GEMM (cluster merge) and of the staged plain warp at 4K.  Outputs land in gpurun_out/.
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:knn|mapper_|warp_|cvt_color|remap_|ransac|gather_matches|chain_" -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --sm-partition 0 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
[ "$1" = list ] && exit 0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn2_tc_kernel -s 3 -c 1 -f -o gpurun_out/knn_tc_cluster \
    python tools/prof_knn.py hamming > gpurun_out/ncu_knn.log 2>&1
echo "knn rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_staged -s 20 -c 1 -f -o gpurun_out/k3s_4k \
    python tools/warp_time.py > gpurun_out/ncu_k3s.log 2>&1
echo "k3s rc=$?"
ls -la gpurun_out/*.ncu-rep
