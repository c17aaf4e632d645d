#!/usr/bin/env bash
# One gpurun call = several bounded stages; every stage logs to gpurun_out/ so a later failure loses nothing.
# usage: tools/gpu_stage.sh stage1 [stage2 ...]      stages: simt probe tests smoke bench bench_c3 ncu_list ncu_full
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD:$PWD/oracle:${PYTHONPATH:-}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/gpu_info.txt 2>&1
for stage in "$@"; do
  echo "=== stage $stage $(date +%T)"
  case "$stage" in
    simt)   DALLE_B200_GEMM=simt timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests_simt.log; tail -15 gpurun_out/tests_simt.log ;;
    simt_all) DALLE_B200_GEMM=simt timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/tests_simt.log; tail -60 gpurun_out/tests_simt.log ;;
    probe)  timeout 900 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; cat gpurun_out/gemm_probe.log ;;
    tests)  timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/tests_gpu.log; tail -60 gpurun_out/tests_gpu.log ;;
    smoke)  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log ;;
    bench)  timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -12 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json ;;
    bench_simt) DALLE_B200_GEMM=simt timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_simt.json 2> gpurun_out/bench_c2_simt.err; tail -3 gpurun_out/bench_c2_simt.err; cat gpurun_out/bench_c2_simt.json ;;
    bench_rows) DALLE_B200_EPI=rows timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_rows.json 2> gpurun_out/bench_c2_rows.err; tail -3 gpurun_out/bench_c2_rows.err ;;
    bench_cols) DALLE_B200_EPI=cols timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_cols.json 2> gpurun_out/bench_c2_cols.err; tail -3 gpurun_out/bench_c2_cols.err ;;
    bench_auto) timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_auto.json 2> gpurun_out/bench_c2_auto.err; tail -3 gpurun_out/bench_c2_auto.err ;;
    bench_c3) timeout 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json ;;
    bench_c4) timeout 900 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json ;;
    bench_ref) timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
    attn_probe) for be in tc mma; do for pat in full axial_row axial_col; do timeout 120 python tools/attn_probe.py --backend $be --pattern $pat; done; done 2>&1 | grep "^\[" | tee gpurun_out/attn_probe.log ;;
    ncu_attn) timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -s 3 -c 4 -o gpurun_out/prof_attn -f python tools/attn_probe.py --backend tc --iters 1 > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log ;;
    ncu_ew) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bwd_tma|scale_bwd_slab|geglu_bwd_kernel" -s 2 -c 4 -o gpurun_out/prof_ew -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --extra "" > gpurun_out/ncu_ew.log 2>&1; tail -2 gpurun_out/ncu_ew.log ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --extra "" > gpurun_out/ncu_list.log 2>&1; tail -3 gpurun_out/ncu_list.log ;;
    ncu_list_c3) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline --no-graph --extra '' > gpurun_out/ncu_list_c3.log 2>&1; tail -3 gpurun_out/ncu_list_c3.log ;;
    attn_probe_tc) for pat in full axial_row axial_col; do timeout 120 python tools/attn_probe.py --backend tc --pattern $pat; done 2>&1 | grep "^\[" | tee gpurun_out/attn_probe.log; for pat in axial_row axial_col; do timeout 120 python tools/attn_probe.py --backend tc --pattern $pat --gather 0; done 2>&1 | grep "^\[" | tee -a gpurun_out/attn_probe.log ;;
    sanitizer) SEL="test_ln_shift_fwd_bwd or (test_gemm_store_all_majors and 96-64-72) or test_gemm_resid_geglu_epilogues or (test_attention_fwd_bwd_patterns and 24-9-4) or (test_axial_gather_kernels and 21-16) or test_dropout_kernel or test_scale_bwd_colsum or test_token_embedding"
      for tool in memcheck racecheck; do
        timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
        echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
      done ;;
    ncu_attn_gather) timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -s 3 -c 4 -o gpurun_out/prof_attn_gather -f python tools/attn_probe.py --backend tc --pattern axial_col --iters 1 > gpurun_out/ncu_attn_gather.log 2>&1; tail -2 gpurun_out/ncu_attn_gather.log ;;
    ncu_full) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 44 -c 20 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --extra "" > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done $(date +%T)"
