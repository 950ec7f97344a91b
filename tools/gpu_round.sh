#!/bin/bash
# One gpurun call for a round's evidence: GPU parity tests, smoke, both bench arms, then (each only
# after the same command exited 0 without ncu) the ncu launch list of bench.py and full captures of
# the dominant kernels.  Logs land in gpurun_out/; copy what is to be judged into profiles/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then echo "tests failed: skipping benches"; exit 0; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/bench_default.log | cut -c1-900
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1
echo "reference arm exit $?"; tail -c 700 gpurun_out/bench_reference.log
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launch_bench.log 2>&1
P="python tools/profile_target.py --engine persistent --segments 127 --iters 2"
$P > gpurun_out/plain_persist.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_persist -s 1 -c 1 -f -o gpurun_out/prof_persist $P > gpurun_out/ncu_persist.log 2>&1
Q="python tools/profile_target.py --beam 128 --segments 8 --iters 1"
$Q > gpurun_out/plain_bs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bs_pass -c 1 -f -o gpurun_out/prof_bs $Q > gpurun_out/ncu_bs.log 2>&1
R="python tools/profile_target.py --engine persistent --segments 8 --iters 2"
$R > gpurun_out/plain_level.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_level16 -c 1 -f -o gpurun_out/prof_level16 $R > gpurun_out/ncu_level16.log 2>&1
FLASHV_TRACE_FILE=gpurun_out/trace.bin $P > /dev/null 2>&1 && python tools/trace_report.py gpurun_out/trace.bin > gpurun_out/phase_trace.txt 2>&1
cat gpurun_out/plain_persist.log gpurun_out/plain_bs.log gpurun_out/plain_level.log
