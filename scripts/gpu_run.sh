#!/bin/bash
# One parameterised GPU script (replaces the per-experiment gpu_round*/gpu_multi* files of round 1).
#   gpurun --timeout 900 -- 'bash scripts/gpu_run.sh <stage> [<stage> ...]'
# Every stage writes under gpurun_out/<tag>/ (tag = $SB_TAG or "run") and prints a short summary.
set -u
TAG=${SB_TAG:-run}
F=gpurun_out/$TAG
mkdir -p "$F"
N=${SB_NGPU:-1}
PORT=29871

pick() { # one summary line from a bench JSON line on stdin
  grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
r=d.get('roofline') or {}
pc=d.get('parity_check') or {}
hx=d.get('halo_exchange')
if hx: print('  halo_exchange us %.1f' % hx['us'], 'ref', hx.get('reference_us'), 'ref_cudampi', hx.get('reference_cudampi_us'), 'ours_cpp', hx.get('ours_cpp_one_process_us'), 'nvlink GB/s/dir %.1f' % hx['nvlink_gbs_per_dir'])
print('$1', 'n', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'kernel %.4f' % r.get('kernel_ms', 0), 'frac %.3f' % r.get('step_frac_of_roofline', 0), 'per_gpu %.4g' % d['per_gpu'], 'launches', d['gpu_launches'], d['config']['schedule'], 'parity', pc.get('bit_exact'), pc.get('schedule'))"
}

bench() { # bench <label> <extra args...>   (N ranks under torchrun when N > 1)
  local label=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus "$N" "$@" 2>"$F/bench_${label}.err" | tee "$F/bench_${label}.json" | pick "$label"
    PORT=$((PORT + 1))
  else
    timeout 400 python bench.py "$@" 2>"$F/bench_${label}.err" | tee "$F/bench_${label}.json" | pick "$label"
  fi
  tail -3 "$F/bench_${label}.err" | cut -c1-300
}

for stage in "$@"; do
  echo "=== stage $stage (N=$N)"
  case $stage in
  tests) # the whole GPU suite
    timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -70 | tee "$F/pytest_gpu.txt" | cut -c1-220 ;;
  tests_jacobi)
    timeout 900 python -m pytest tests/test_gpu_jacobi.py -q -m gpu -x 2>&1 | tail -40 | tee "$F/pytest_jacobi.txt" | cut -c1-250 ;;
  tests_mp) # multi-rank / multi-GPU parity (python one-process-per-GPU + in-process, C++ ranks under sb_mpirun)
    timeout 900 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_cpp_api.py -q -m gpu -x -k "one_process or multi_gpu or multi_rank or one_rank_per_gpu" 2>&1 | tail -8 | tee "$F/pytest_mp.txt" ;;
  mp_check) # the torchrun parity script itself, with its per-shape lines
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29861 tests/mp_exchange_check.py 2>"$F/mp_check.err" | tee "$F/mp_check.txt" | cut -c1-400
    tail -5 "$F/mp_check.err" | cut -c1-400 ;;
  smoke)
    timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee "$F/smoke.txt" ;;
  bench) # the line the driver reads
    bench default --steps 30 --warmup 5 ;;
  bench_quick)
    bench quick --steps 30 --warmup 5 --no-cpu-baseline --no-e2e ;;
  bench_f32)
    bench f32 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --dtype f32 ;;
  bench_queued)
    bench queued --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --schedule queued --no-parity ;;
  bench_diag) # where the multi-rank overhead of the fused kernel goes (timing only: some of these compute wrong halos)
    bench again --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    SB_DEBUG_FUSED=1 bench nowait --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    SB_DEBUG_FUSED=3 bench nowait_nosignal --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    SB_DEBUG_NOPUSH=1 bench nopush --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity ;;
  bench_diag2)
    SB_DEBUG_FUSED=3 bench nowait_nosignal --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    SB_DEBUG_NOPUSH=1 bench nopush --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity ;;
  bench_cuts) # which face direction costs what between ranks (2 ranks: one cut along y, then z)
    for c in y z; do
      SB_BENCH_CUT=$c bench "cut${c}" --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
      SB_BENCH_CUT=$c SB_DEBUG_FUSED=3 bench "cut${c}_nowait_nosignal" --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
      SB_BENCH_CUT=$c SB_DEBUG_NOPUSH=1 bench "cut${c}_nopush" --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    done ;;
  bench_sig) # what a signal costs: the real release, a flag without fence, a 2.5 us sleep without fence (2 ranks, cut along y)
    for v in 0 4 8; do
      SB_BENCH_CUT=y SB_DEBUG_FUSED=$v bench "sig$v" --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    done
    SB_DEBUG_FUSED=4 bench "sig4_xcut" --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity ;;
  bench_rot) # the rotated walk over the tiles, on and off
    SB_FUSED_ROTATE=0 bench rot0 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity
    SB_FUSED_ROTATE=1 bench rot1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity ;;
  time_f32) # FP32 fused kernel variants on one GPU
    DTYPE=f32 F32VARIANTS=1 timeout 600 python scripts/time_fused.py 512 30 2>&1 | tail -14 | tee "$F/time_f32.txt" ;;
  time_fused) # kernel variants + single-GPU stand-ins for the multi-rank kernels, one box
    timeout 900 python scripts/time_fused.py 512 30 2>&1 | tail -24 | tee "$F/time_fused.txt" ;;
  ncu_fused2) # the fused kernel of the 2-subdomain stand-in (half of its CTAs are boundary CTAs) under ncu
    ONLY=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_fused -s 12 -c 2 -o "$F/prof_jacobi_fused2" -f python scripts/time_fused.py 512 4 >"$F/ncu_fused2.log" 2>&1
    tail -3 "$F/ncu_fused2.log" ;;
  debug_astaroth)
    CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_astaroth.py -q -m gpu -x -k "iteration_through" 2>&1 | tail -70 | cut -c1-220
    timeout 600 python -m pytest tests/test_gpu_astaroth.py -q -m gpu -x 2>&1 | tail -70 | cut -c1-220 ;;
  ncu_nvlink) # NVLink byte counters of the exchange kernel and of the fused jacobi kernel, 1 process x 2 GPUs (needs >= 2 GPUs)
    timeout 600 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:box_copy -s 10 -c 8 --csv --log-file "$F/nvlink_box_copy.csv" python scripts/time_exchange_mg.py 256 2 3 float32 >"$F/nvlink_box_copy.log" 2>&1
    tail -9 "$F/nvlink_box_copy.csv" | cut -c1-260
    NGPU=2 ONLY=2 timeout 600 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:jacobi_fused -s 8 -c 4 --csv --log-file "$F/nvlink_fused.csv" python scripts/time_fused.py 512 4 >"$F/nvlink_fused.log" 2>&1
    tail -5 "$F/nvlink_fused.csv" | cut -c1-260 ;;
  tma_pack) # TMA 3-D pack experiment vs the LSU path (scripts/exp/tma_pack3d.cu, built in the container)
    timeout 120 scripts/exp/tma_pack3d 2>&1 | tail -12 | tee "$F/tma_pack3d.txt" ;;
  tests_astaroth)
    timeout 900 python -m pytest tests/test_gpu_astaroth.py -q -m gpu -x 2>&1 | tail -60 | tee "$F/pytest_astaroth.txt" | cut -c1-250 ;;
  bench_launchsync) # the round-1 handshake (separate wait / signal launches) for the before/after
    SB_FUSED_INKERNEL=0 bench launchsync --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity ;;
  bench_reference)
    timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee "$F/bench_reference.json" | cut -c1-400 ;;
  launches) # every launch of the bench command, cold and serialised: compare SHARES
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file "$F/launches.csv" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity >"$F/ncu_launches.log" 2>&1
    tail -12 "$F/launches.csv" | cut -c1-200 ;;
  ncu_fused)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_fused_kernel -s 4 -c 1 -o "$F/prof_jacobi_fused" -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity >"$F/ncu_full.log" 2>&1
    tail -3 "$F/ncu_full.log" ;;
  ncu_fused_f32)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_fused_kernel -s 4 -c 1 -o "$F/prof_jacobi_fused_f32" -f python bench.py --dtype f32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-exchange-bench --no-parity >"$F/ncu_full_f32.log" 2>&1
    tail -3 "$F/ncu_full_f32.log" ;;
  ncu_astaroth) # the astaroth substep kernels, FP64 256^3 (variant in SB_AC_VARIANTS, default "team tile")
    SKIP_CELL=1 SKIP_ITER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ac_t -c 6 -o "$F/prof_astaroth_f64" -f python scripts/time_astaroth.py 256 f64 1 >"$F/ncu_astaroth.log" 2>&1
    tail -3 "$F/ncu_astaroth.log" ;;
  sanitize) # memcheck on the small jacobi cases (plain, regions, fused, in-process multi-subdomain)
    timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_jacobi.py -q -m gpu -k "not full_size and not golden" >"$F/sanitize_full.txt" 2>&1
    grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" "$F/sanitize_full.txt" | sort | uniq -c | sort -rn | head -12 | tee "$F/sanitize.txt"
    grep -E "Invalid" -A 12 "$F/sanitize_full.txt" | head -40 | cut -c1-200 ;;
  cpp) # the C++ API: the reference's own suites and drivers against our library
    ( timeout 600 bin/test_cuda 2>&1 | tail -3 ) | tee "$F/test_cuda.txt"
    ( timeout 300 bin/test_cpu 2>&1 | tail -3 ) | tee "$F/test_cpu.txt" ;;
  golden) # reference-generated golden vectors (oracle/ref/*.py), written to gpurun_out/
    timeout 300 python oracle/ref/make_jacobi_golden.py 2>&1 | tail -8 ;;
  astaroth)
    SKIP_CELL=1 timeout 300 python scripts/time_astaroth.py 256 f64 5 2>&1 | tail -16 | tee "$F/astaroth_f64.txt"
    SKIP_CELL=1 timeout 300 python scripts/time_astaroth.py 256 f32 5 2>&1 | tail -10 | tee "$F/astaroth_f32.txt" ;;
  *) # anything else: a script path with arguments in SB_ARGS
    if [ -f "$stage" ]; then timeout 900 bash "$stage" 2>&1 | tail -40 | tee "$F/$(basename "$stage").txt"; else echo "unknown stage $stage"; fi ;;
  esac
done
rm -f plan_*.txt gpurun_out/plan_*.txt
