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
    prio)    { for a in "7 1" "7 0" "4 1"; do echo "== competitor CTAs/SM, ilp: $a"; timeout 120 _variants/prio $a; done; } > $OUT/prio.log 2>&1; cat $OUT/prio.log ;;
    edge)    timeout 900 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -s -rf > $OUT/edge.log 2>&1; grep -v "^$" $OUT/edge.log | tail -25 ;;
    stats3)  { for sc in bvh_heavy rtcamp6; do HNM_TRACE_STATS=1 timeout 300 python tools/time_passes.py $sc 1920 1080 6; done; } > $OUT/stats3.log 2>&1; cat $OUT/stats3.log ;;
    parity)  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_multi.py -m gpu -q -x -rf > $OUT/parity.log 2>&1; tail -8 $OUT/parity.log ;;
    corun)   { bash tools/ab.sh "HNM_X=0" "HNM_PROFILE_OVERLAP=1" "HNM_RNG_OVERLAP=0"; for sc in diamond; do HNM_PROFILE_OVERLAP=1 timeout 200 python tools/time_passes.py $sc 1920 1080 9 2>&1 | tail -3; done; } > $OUT/corun.log 2>&1; cat $OUT/corun.log ;;
    bench2s) timeout 600 python bench.py --config 2 --steps 8 --no-e2e --no-cpu --no-traffic > $OUT/bench2s.json 2> $OUT/bench2s.err; python -c "
import json,sys
d=json.loads(open('$OUT/bench2s.json').read().strip().splitlines()[-1]); print('config2 value', d['value'], 'ms/pass', d['ms_per_step']/d['config']['passes_per_step'], {k: round(v/d['detail']['profiled_passes'],3) for k,v in d['detail']['kernel_ms'].items()})" ;;
    bench4s) timeout 600 python bench.py --config 4 --steps 2 --no-e2e --no-cpu --no-traffic > $OUT/bench4s.json 2> $OUT/bench4s.err; python -c "
import json,sys
d=json.loads(open('$OUT/bench4s.json').read().strip().splitlines()[-1]); print('config4 value', d['value'], 'ms/pass', d['ms_per_step']/d['config']['passes_per_step'])" ;;
    bench3s) timeout 600 python bench.py --config 3 --steps 4 --no-e2e --no-cpu --no-traffic > $OUT/bench3s.json 2> $OUT/bench3s.err; python -c "
import json,sys
d=json.loads(open('$OUT/bench3s.json').read().strip().splitlines()[-1]); print('config3 value', d['value'], 'ms/pass', d['ms_per_step']/d['config']['passes_per_step'], {k: round(v/d['detail']['profiled_passes'],3) for k,v in d['detail']['kernel_ms'].items()})" ;;
    bisect)  { for lib in "" _variants/nodyn.so _variants/nomiss.so _variants/nosurf.so _variants/noneer.so _variants/noconf.so; do
                 HNM_CORE_LIB=$lib timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -2; done
               HNM_RNG_OVERLAP=0 timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -2
               timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 1 2>&1 | tail -2
               timeout 200 python tools/diag_scene.py tbf3_pl 160 90 1 2 2>&1 | tail -2; } > $OUT/bisect.log 2>&1; cat $OUT/bisect.log ;;
    race)    { timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/diag_scene.py rtcamp5_pl 64 36 1 1 2>&1 | grep -v "^$" | head -80;
               timeout 600 compute-sanitizer --tool synccheck python tools/diag_scene.py rtcamp5_pl 64 36 1 1 2>&1 | grep -v "^$" | head -40; } > $OUT/race.log 2>&1; head -c 7000 $OUT/race.log ;;
    bisect2) { for lib in _variants/surf2.so _variants/surf1nosort.so _variants/surf2nosort.so; do
                 HNM_CORE_LIB=$lib timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -2; done; } > $OUT/bisect2.log 2>&1; cat $OUT/bisect2.log ;;
    bisect3) { export HNM_DIAG_QUEUES=1; for lib in _variants/nosurf.so ""; do
                 HNM_CORE_LIB=$lib HNM_RNG_OVERLAP=0 timeout 200 python tools/diag_scene.py rtcamp5_pl 64 36 1 1 2>&1 | tail -4; done; } > $OUT/bisect3.log 2>&1; cat $OUT/bisect3.log ;;
    bisect4) { for lib in _variants/dynO1.so _variants/dyninl.so _variants/dyn3blk.so; do
                 HNM_CORE_LIB=$lib HNM_RNG_OVERLAP=0 timeout 200 python tools/diag_scene.py rtcamp5_pl 64 36 1 1 2>&1 | tail -2; done;
               HNM_RNG_OVERLAP=0 timeout 300 compute-sanitizer --tool initcheck python tools/diag_scene.py rtcamp5_pl 64 36 1 1 2>&1 | grep -v "^$" | head -40; } > $OUT/bisect4.log 2>&1; head -c 5000 $OUT/bisect4.log ;;
    abdyn)   { for lib in _variants/nodyn.so _variants/nosurf.so _variants/dynblk3.so _variants/dyn3blk.so; do
                 echo "== $lib"; HNM_CORE_LIB=$lib timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -1 | cut -c1-80
                 HNM_CORE_LIB=$lib timeout 300 python bench.py --config 2 --steps 8 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config2', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3), {k: round(v/d['detail']['profiled_passes'],3) for k,v in d['detail']['kernel_ms'].items() if k in ('confirm','shade_miss','shade_delta','shade_nee','nee_resolve')})"
                 HNM_CORE_LIB=$lib timeout 300 python bench.py --config 4 --steps 2 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config4', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"
               done; } > $OUT/abdyn.log 2>&1; cat $OUT/abdyn.log ;;
    fast)    { timeout 600 python -m pytest tests/test_gpu_reference_chain.py -k fast_math -m gpu -q -s 2>&1 | grep "FAST_MATH\|passed\|failed";
               for prec in exact fast; do for c in 2 4; do timeout 300 python bench.py --config $c --steps 4 --precision $prec --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$prec config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3), {k: round(v/d['detail']['profiled_passes'],3) for k,v in d['detail']['kernel_ms'].items() if k in ('shade_miss','shade_delta','shade_nee','nee_resolve')})"; done; done; } > $OUT/fast.log 2>&1; cat $OUT/fast.log ;;
    edge2)   timeout 900 python -m pytest tests/test_gpu_edge_cases.py -k directed -m gpu -q -s -rf 2>&1 | grep "directed rays\|passed\|failed" > $OUT/edge2.log; cat $OUT/edge2.log ;;
    tma)     { for sc in rtcamp6 bvh_heavy; do for e in 0 1; do echo "== $sc HNM_CONFIRM_TMA=$e";
                 HNM_CONFIRM_TMA=$e timeout 200 python tools/diag_scene.py $sc 160 90 1 2 2>&1 | tail -1 | cut -c1-90
                 HNM_CONFIRM_TMA=$e HNM_RNG_OVERLAP=0 timeout 300 python tools/time_passes.py $sc 1920 1080 6 2>&1 | tail -2; done; done; } > $OUT/tma.log 2>&1; cat $OUT/tma.log ;;
    prof)    RND=r02
             timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$RND.csv python bench.py --steps 2 --warmup 3 --pps 1 --no-e2e --no-cpu --no-traffic > $OUT/prof_bench.log 2>&1
             export HNM_RNG_OVERLAP=0
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -f -o gpurun_out/prof_trace_$RND python tools/traffic_probe.py rtcamp6 1920 1080 > $OUT/p1.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -f -o gpurun_out/prof_trace3_$RND python tools/traffic_probe.py bvh_heavy 1920 1080 > $OUT/p2.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_confirm_pairs -c 2 -f -o gpurun_out/prof_confirm_$RND python tools/traffic_probe.py rtcamp6 1920 1080 > $OUT/p3.log 2>&1
             HNM_CONFIRM_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_confirm -c 2 -f -o gpurun_out/prof_confirmtma_$RND python tools/traffic_probe.py rtcamp6 1920 1080 > $OUT/p4.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_isaac_raygen_tm -c 1 -f -o gpurun_out/prof_isaac_$RND python tools/traffic_probe.py rtcamp6 1920 1080 > $OUT/p5.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade_surf -c 2 -f -o gpurun_out/prof_shade_$RND python tools/traffic_probe.py rtcamp6 1920 1080 > $OUT/p6.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nee_resolve -c 1 -f -o gpurun_out/prof_neer_$RND python tools/traffic_probe.py rtcamp6 1920 1080 > $OUT/p7.log 2>&1
             ls -la gpurun_out/*$RND* ;;
    young)   { for e in "HNM_X=0" "HNM_ISAAC_ROUNDS=1" "HNM_ISAAC_ROUNDS=1 HNM_RNG_START_BOUNCE=0" \
                 "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7" \
                 "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7 HNM_ISAAC_ROUNDS=1" \
                 "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7 HNM_ISAAC_ROUNDS=1 HNM_RNG_START_BOUNCE=0" \
                 "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=7 HNM_ISAAC_ROUNDS=4 HNM_RNG_START_BOUNCE=0" \
                 "HNM_CORE_LIB=_variants/isaac64r.so HNM_TRACE_BLOCKS=6 HNM_ISAAC_ROUNDS=1 HNM_RNG_START_BOUNCE=0"; do
                 echo "== $e"
                 for c in 2 4; do env $e timeout 300 python bench.py --config $c --steps 6 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/young.log 2>&1; cat $OUT/young.log ;;
    tm1)     { timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "isaac or config1 or rng" 2>&1 | tail -5
               bash tools/ab.sh "HNM_ISAAC_TMEM=0" "HNM_ISAAC_TMEM=1" "HNM_ISAAC_TMEM=1 HNM_RNG_OVERLAP=0" "HNM_ISAAC_TMEM=1 HNM_TRACE_BLOCKS=6" "HNM_ISAAC_TMEM=1 HNM_TRACE_BLOCKS=7"; } > $OUT/tm1.log 2>&1; cat $OUT/tm1.log ;;
    tm2)     { for v in tmdiag1 tmdiag2 tmdiag3; do echo "== $v"; HNM_CORE_LIB=_variants/$v.so HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60; done
               HNM_RNG_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_isaac_raygen -c 1 -f -o gpurun_out/prof_isaactm python tools/traffic_probe.py rtcamp6 1920 1080 2>&1 | tail -2; } > $OUT/tm2.log 2>&1; cat $OUT/tm2.log ;;
    tm3)     { timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "isaac or config1 or rng" 2>&1 | tail -3
               for v in tmdiag2 tmnoopq; do echo "== $v"; HNM_CORE_LIB=_variants/$v.so HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60; done
               bash tools/ab.sh "HNM_ISAAC_TMEM=0 HNM_RNG_OVERLAP=0" "HNM_ISAAC_TMEM=1 HNM_RNG_OVERLAP=0" "HNM_ISAAC_TMEM=0" "HNM_ISAAC_TMEM=1" "HNM_ISAAC_TMEM=1 HNM_TRACE_BLOCKS=6"; } > $OUT/tm3.log 2>&1; cat $OUT/tm3.log ;;
    tm4)     { HNM_CORE_LIB=_variants/tmdiag2.so HNM_RNG_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_isaac_raygen -c 1 -f -o gpurun_out/prof_isaactm_d2 python tools/traffic_probe.py rtcamp6 1920 1080 2>&1 | tail -2; } > $OUT/tm4.log 2>&1; cat $OUT/tm4.log ;;
    tm5)     { for e in "HNM_ISAAC_TMEM=0" "HNM_ISAAC_TMEM=1" "HNM_ISAAC_TMEM=1 HNM_RNG_OVERLAP=0" "HNM_ISAAC_TMEM=1 HNM_RNG_START_BOUNCE=0" "HNM_ISAAC_TMEM=1 HNM_RNG_START_BOUNCE=2" "HNM_ISAAC_TMEM=1 HNM_TRACE_BLOCKS=6"; do
                 echo "== $e"
                 for c in 2 4 3; do env $e timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/tm5.log 2>&1; cat $OUT/tm5.log ;;
    tm6)     { for v in tmhi; do echo "== $v"; HNM_CORE_LIB=_variants/$v.so timeout 200 python tools/diag_scene.py rtcamp6 160 90 1 2 2>&1 | tail -1 | cut -c1-100; HNM_CORE_LIB=_variants/$v.so HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60; done; } > $OUT/tm6.log 2>&1; cat $OUT/tm6.log ;;
    tm7)     { for v in tmdiag3 tmdiag4 tmstag; do echo "== $v"; HNM_CORE_LIB=_variants/$v.so HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60; done; } > $OUT/tm7.log 2>&1; cat $OUT/tm7.log ;;
    tm8)     { timeout 200 python tools/diag_scene.py rtcamp6 160 90 1 2 2>&1 | tail -1 | cut -c1-100; bash tools/ab.sh "HNM_RNG_OVERLAP=0" "HNM_X=1"; } > $OUT/tm8.log 2>&1; cat $OUT/tm8.log ;;
    tm9)     { timeout 200 python tools/diag_scene.py rtcamp6 160 90 1 2 2>&1 | tail -1 | cut -c1-100; HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60
               for v in tmcs0 tmcs64 tmcs256; do echo "== $v"; HNM_CORE_LIB=_variants/$v.so HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60; done; } > $OUT/tm9.log 2>&1; cat $OUT/tm9.log ;;
    tm10)    { for e in "HNM_X=1" "HNM_RNG_OVERLAP=0" "HNM_CORE_LIB=_variants/tm64.so" "HNM_CORE_LIB=_variants/tm64.so HNM_TRACE_BLOCKS=6"; do
                 echo "== $e"
                 for c in 2 4; do env $e timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/tm10.log 2>&1; cat $OUT/tm10.log ;;
    sl1)     { timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "isaac or config1 or rng or batching or scenes_bit_exact" 2>&1 | tail -3
               for e in "HNM_RNG_SLICES=0" "HNM_RNG_SLICES=1" "HNM_RNG_SLICES=2" "HNM_RNG_SLICES=3" "HNM_RNG_SLICES=5"; do
                 echo "== $e"
                 for c in 2 4; do env $e timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/sl1.log 2>&1; cat $OUT/sl1.log ;;
    sl2)     { for e in "HNM_RNG_SLICES=5 HNM_PROFILE_OVERLAP=1" "HNM_RNG_SLICES=0 HNM_PROFILE_OVERLAP=1" "HNM_RNG_OVERLAP=0"; do echo "== $e"; env $e timeout 200 python tools/time_passes.py rtcamp6 1920 1080 12 2>&1 | tail -2; done; } > $OUT/sl2.log 2>&1; cat $OUT/sl2.log ;;
    bt1)     { for b in 3 4 6 8; do
                 echo "== batch $b"
                 for c in 2 4; do timeout 300 python bench.py --config $c --steps 4 --batch $b --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/bt1.log 2>&1; cat $OUT/bt1.log ;;
    sh1)     { timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -1 | cut -c1-100; HNM_SHADE_THREADS=128 timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -1 | cut -c1-100
               for e in "HNM_X=1" "HNM_SHADE_THREADS=128" "HNM_SHADE_THREADS=64"; do
                 echo "== $e"
                 for c in 2 4; do env $e timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3), {k: round(v/d['detail']['profiled_passes'],3) for k,v in d['detail']['kernel_ms'].items()})"; done; done; } > $OUT/sh1.log 2>&1; cat $OUT/sh1.log ;;
    cf1)     { timeout 200 python tools/setup_time.py 2>&1 | tail -2
               HNM_RNG_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_confirm -c 2 -f -o gpurun_out/prof_confirm3 python tools/traffic_probe.py bvh_heavy 1920 1080 2>&1 | tail -1
               HNM_RNG_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nee_resolve -c 1 -f -o gpurun_out/prof_neer3 python tools/traffic_probe.py bvh_heavy 1920 1080 2>&1 | tail -1; } > $OUT/cf1.log 2>&1; cat $OUT/cf1.log ;;
    cp1)     { timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "scenes_bit_exact or config1 or overflow or ties or full_size" 2>&1 | tail -3
               for sc in rtcamp6 bvh_heavy diamond; do for e in 0 1; do echo "== $sc HNM_CONFIRM_PAIRS=$e"; HNM_CONFIRM_PAIRS=$e HNM_RNG_OVERLAP=0 timeout 300 python tools/time_passes.py $sc 1920 1080 6 2>&1 | tail -1 | cut -c1-200; done; done; } > $OUT/cp1.log 2>&1; cat $OUT/cp1.log ;;
    cp2)     { HNM_RNG_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_confirm_pairs -c 3 -f -o gpurun_out/prof_cpairs3 python tools/traffic_probe.py bvh_heavy 1920 1080 2>&1 | tail -1
               HNM_TRACE_STATS=1 HNM_RNG_OVERLAP=0 timeout 300 python tools/time_passes.py bvh_heavy 1920 1080 3 2>&1 | tail -2; } > $OUT/cp2.log 2>&1; cat $OUT/cp2.log ;;
    cp3)     { for lib in "" _variants/cand16.so _variants/cand32.so; do for sc in bvh_heavy rtcamp6; do echo "== $sc lib=$lib"; HNM_CORE_LIB=$lib timeout 200 python tools/diag_scene.py $sc 160 90 1 2 2>&1 | tail -1 | cut -c1-90; HNM_CORE_LIB=$lib HNM_TRACE_STATS=1 HNM_RNG_OVERLAP=0 timeout 300 python tools/time_passes.py $sc 1920 1080 3 2>&1 | tail -2; done; done; } > $OUT/cp3.log 2>&1; cat $OUT/cp3.log ;;
    sah1)    { for e in "HNM_X=1" "HNM_SAH_MAXLEAF=4" "HNM_SAH_MAXLEAF=12" "HNM_SAH_CT=0.5" "HNM_SAH_CT=2.0" "HNM_SAH_CT=0.5 HNM_SAH_MAXLEAF=4" "HNM_SAH_CT=2.0 HNM_SAH_MAXLEAF=12"; do for sc in bvh_heavy; do echo "== $sc $e"; env $e HNM_TRACE_STATS=1 HNM_RNG_OVERLAP=0 timeout 300 python tools/time_passes.py $sc 1920 1080 3 2>&1 | tail -2 | cut -c1-230; done; done; } > $OUT/sah1.log 2>&1; cat $OUT/sah1.log ;;
    c1a)     { for e in "HNM_X=1" "HNM_RNG_SLICES=0" "HNM_RNG_SLICES=1" "HNM_RNG_OVERLAP=0"; do echo "== $e"; env $e timeout 300 python bench.py --config 1 --steps 64 --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config 1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3))"; done; } > $OUT/c1a.log 2>&1; cat $OUT/c1a.log ;;
    cx1)     { timeout 200 python tools/diag_scene.py rtcamp6 160 90 1 2 2>&1 | tail -1 | cut -c1-100
               for lib in "" _variants/tmx32.so; do echo "== lib=$lib"; HNM_CORE_LIB=$lib HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-60
                 for c in 2 4; do HNM_CORE_LIB=$lib timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/cx1.log 2>&1; cat $OUT/cx1.log ;;
    final)   { timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
               timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -6; } > $OUT/final.log 2>&1; cat $OUT/final.log ;;
    so1)     { for lib in ${SO_LIBS:-"" _variants/sort0.so _variants/sort2.so}; do echo "== lib=$lib"; HNM_CORE_LIB=$lib timeout 200 python tools/diag_scene.py rtcamp5_pl 160 90 1 2 2>&1 | tail -1 | cut -c1-100; HNM_CORE_LIB=$lib HNM_RNG_OVERLAP=0 timeout 200 python tools/time_passes.py rtcamp6 1920 1080 3 2>&1 | tail -1 | cut -c1-200
                 for c in 2 4; do HNM_CORE_LIB=$lib timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/so1.log 2>&1; cat $OUT/so1.log ;;
    e2e1)    { timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_halves or resolve" 2>&1 | tail -2
               timeout 600 python bench.py --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('config 2 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['seconds'])"; } > $OUT/e2e1.log 2>&1; cat $OUT/e2e1.log ;;
    sl3)     { for e in "HNM_RNG_SLICES=4" "HNM_RNG_SLICES=6" "HNM_RNG_SLICES=8"; do echo "== $e"; for c in 2 4 3; do env $e timeout 300 python bench.py --config $c --steps 4 --no-e2e --no-cpu --no-traffic 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  config $c', round(d['value'],1), 'ms/pass', round(d['ms_per_step']/d['config']['passes_per_step'],3))"; done; done; } > $OUT/sl3.log 2>&1; cat $OUT/sl3.log ;;
    race2)   { timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -4
               timeout 400 compute-sanitizer --tool synccheck python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -3; } > $OUT/race2.log 2>&1; cat $OUT/race2.log ;;
    sync2)   { timeout 400 compute-sanitizer --tool synccheck --print-limit 6 python tools/sanitize_run.py 2>&1 | grep -v "^$" | head -60; } > $OUT/sync2.log 2>&1; cat $OUT/sync2.log ;;
    *) echo "unknown step $step" ;;
  esac
done
