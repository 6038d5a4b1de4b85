# the measurement set copied into profiles/ (run on the GPU box):  bash tools/probes/final_profile.sh <tag> a|b
#   a: bench line, reference arm, per-workload lines (C1, C3 at 1 GiB, C5's single-GPU share), ncu launch list
#   b: ncu --set full of the top kernels + summary  (one profiler pass per call)
#   c: (back in the development container, no GPU) regenerate profiles/traffic.json from the capture of b and copy the set into profiles/
set -x
T=${1:-r02z}
case ${2:-a} in
a)
  timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
  timeout 300 python bench.py --workload C1 --steps 20 --warmup 3 > gpurun_out/${T}_bench_C1.json 2> gpurun_out/${T}_c1.err
  timeout 900 python bench.py --workload C3 --steps 3 --warmup 3 > gpurun_out/${T}_bench_C3.json 2> gpurun_out/${T}_c3.err
  timeout 600 python bench.py --workload C4 --steps 20 --warmup 3 > gpurun_out/${T}_bench_C4.json 2> gpurun_out/${T}_c4.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu1.log 2>&1
  tail -c 700 gpurun_out/${T}_bench.json
  ;;
b)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seq_t|k_exec2|k_huf|k_xxh" -c 4 -o gpurun_out/${T}_top -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu2.log 2>&1
  python tools/ncu_summary.py gpurun_out/${T}_top.ncu-rep --hot 40 > gpurun_out/${T}_top_kernels_ncu_full_summary.txt 2>&1
  head -30 gpurun_out/${T}_top_kernels_ncu_full_summary.txt
  ;;
c)
  python tools/update_traffic.py gpurun_out/${T}_top.ncu-rep --workload C2 --frames 4096 --captured ${T} | tail -3
  cp gpurun_out/${T}_bench.json gpurun_out/${T}_bench_reference_arm.json gpurun_out/${T}_bench_C1.json gpurun_out/${T}_bench_C3.json gpurun_out/${T}_bench_C4.json \
     gpurun_out/${T}_launches.csv gpurun_out/${T}_top_kernels_ncu_full_summary.txt profiles/
  echo "the bench line of step a was printed before traffic.json was regenerated: run python bench.py once more for a line with roofline.traffic"
  ;;
esac
