# the measurement set copied into profiles/ (run on the GPU box)
set -x
T=${1:-r01f}
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seq|k_exec2|k_huf|k_xxh" -c 5 -o gpurun_out/${T}_top -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_top.ncu-rep --hot 40 > gpurun_out/${T}_top_kernels_ncu_full_summary.txt 2>&1
tail -c 600 gpurun_out/${T}_bench.json
