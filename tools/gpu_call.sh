#!/bin/bash
# One gpurun call: everything writes under gpurun_out/ (merged back).  Usage: tools/gpu_call.sh TAG step...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
for step in "$@"; do
  case $step in
    tests)   timeout 1500 python -m pytest tests -m gpu -q -rf --durations=15 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -25 $OUT/pytest.log ;;
    tests_x) timeout 1500 python -m pytest tests -m gpu -q -x -rf > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -25 $OUT/pytest.log ;;
    golden)  HNM_GOLDEN_REPORT=$OUT/golden.json timeout 600 python -m pytest tests/test_reference_golden.py -m gpu -q -s > $OUT/golden.log 2>&1; tail -5 $OUT/golden.log ;;
    bench2)  timeout 900 python bench.py --config 2 > $OUT/bench2.json 2> $OUT/bench2.err; tail -c 3000 $OUT/bench2.json ;;
    bench3)  timeout 900 python bench.py --config 3 --no-traffic > $OUT/bench3.json 2> $OUT/bench3.err; tail -c 2500 $OUT/bench3.json ;;
    bench4)  timeout 900 python bench.py --config 4 --no-traffic > $OUT/bench4.json 2> $OUT/bench4.err; tail -c 2500 $OUT/bench4.json ;;
    bench1)  timeout 600 python bench.py --config 1 --steps 64 --no-traffic > $OUT/bench1.json 2> $OUT/bench1.err; tail -c 1500 $OUT/bench1.json ;;
    ref2)    timeout 900 python bench.py --impl reference --config 2 --steps 3 --warmup 1 > $OUT/ref2.json 2> $OUT/ref2.err; tail -c 1200 $OUT/ref2.json ;;
    ab_wid)
      {
      export HNM_WID_STATS=1
      bash tools/ab.sh "HNM_X=0" "HNM_TRACE_BLOCKS=7" "HNM_RNG_START_BOUNCE=0" "HNM_RNG_OVERLAP=0" \
         "HNM_CORE_LIB=_variants/isaac64r.so" "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7 HNM_RNG_MIDTRACE=1" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7 HNM_RNG_MIDTRACE=1 HNM_RNG_START_BOUNCE=2" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=6 HNM_RNG_MIDTRACE=1"
      } > $OUT/ab_wid.log 2>&1; cat $OUT/ab_wid.log ;;
    diag)    for sc in rtcamp6 diamond; do timeout 300 python tools/diag_directed.py $sc; done > $OUT/diag.log 2>&1; head -c 6000 $OUT/diag.log ;;
    ab_prio)
      {
      export HNM_WID_STATS=1
      bash tools/ab.sh "HNM_X=0" "HNM_CARVEOUT=100" "HNM_CARVEOUT=100 HNM_TRACE_BLOCKS=7" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_CARVEOUT=100 HNM_TRACE_BLOCKS=7" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_CARVEOUT=100 HNM_TRACE_BLOCKS=7 HNM_RNG_MIDTRACE=1" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_CARVEOUT=100 HNM_TRACE_BLOCKS=6 HNM_RNG_MIDTRACE=1" \
         "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7 HNM_RNG_MIDTRACE=1"
      for sc in diamond; do
        for e in "HNM_X=0" "HNM_CORE_LIB=_variants/isaac64r.so HNM_CARVEOUT=100 HNM_TRACE_BLOCKS=7 HNM_RNG_MIDTRACE=1"; do
          echo "== $sc $e"; env $e timeout 200 python tools/time_passes.py $sc 1920 1080 8 2>&1 | tail -4
        done
      done
      } > $OUT/ab_prio.log 2>&1; cat $OUT/ab_prio.log ;;
    golden_s) HNM_GOLDEN_REPORT=$OUT/golden.json timeout 900 python -m pytest tests/test_reference_golden.py tests/test_gpu_reference_chain.py -m gpu -q -s > $OUT/golden.log 2>&1; grep -v "^$" $OUT/golden.log | tail -30 ;;
    multi)   timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rfs > $OUT/multi.log 2>&1; tail -15 $OUT/multi.log ;;
    *) echo "unknown step $step" ;;
  esac
done
