MRTX_LIB=moonrtx_b200/_variants/libmoonb200_stats.so VARIANTS='beam=0,0+ceiling=0;beam=1,2+ceiling=2' python tools/bench_trace.py cfg3 16 2>&1 | grep '"spp"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'ms' not in d: continue
    print(d['kernel'], 'ms', d['ms'])
    print('  primary walk: warp-iters', d['trav_steps'], 'lane-iters', d['trav_step_lanes'], 'eff', d['trav_step_lanes'] / 32 / max(1, d['trav_steps']))
    print('  shadow  walk: warp-iters', d['start_phases'], 'lane-iters', d['start_phase_lanes'], 'eff', d['start_phase_lanes'] / 32 / max(1, d['start_phases']))
    print('  test phases', d['test_phases'], 'lanes', d['test_phase_lanes'], 'per phase', d['test_phase_lanes'] / max(1, d['test_phases']))
    np_ = d['defer']['shadow_reasons'].get('15', 0)
    print('  walk phases', np_, 'alive lanes at start', d['refills'], 'per phase', d['refills'] / max(1, np_))
"
