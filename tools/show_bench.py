"""print the interesting numbers of a bench.py JSON line (scratch helper):  python tools/show_bench.py file.json"""
import json
import sys


def show(b, ind=''):
    if 'error' in b:
        print(ind, b['config']['workload'], 'ERROR', b['error'][:300])
        return
    e = b.get('e2e') or {}
    c = b['config']
    print(ind, c['workload'], 'value %.3e %s  ms %.4f' % (b['value'], b['unit'], b['ms_per_step']),
          '| single-stream %.4f' % c['single_stream_ms_per_step'] if 'single_stream_ms_per_step' in c else '',
          '| e2e %.3e (%.3f ms)' % (e.get('value', 0), e.get('ms_per_step', 0)) if e else '',
          '| serial %.3f' % e['serial']['ms_per_step'] if e.get('serial') else '', '| gather_verified %s' % b['gather_verified'] if 'gather_verified' in b else '')
    if 'host_issue_ms_per_step' in c:
        print(ind, '   host issue %.4f ms/step' % c['host_issue_ms_per_step'], '| by rank', c.get('ms_per_step_by_rank'), c.get('sm_mhz_by_rank'), '| other L2 method %.4f' % c['l2_other_method']['ms_per_step'] if 'l2_other_method' in c else '')
    if 'stage_ms' in b:
        print(ind, '   stages', {k: round(v, 4) for k, v in b['stage_ms'].items()})
    if 'launches_per_step' in b:
        print(ind, '   launches', b['launches_per_step']['kernel_nodes'], '/', b['launches_per_step']['graph_nodes'], 'graph nodes')
    if 'cpu_baseline' in b:
        cb = b['cpu_baseline']
        print(ind, '   cpu %.3e (%s cores)' % (cb['value'], cb.get('cores')), 'one thread: %.3e' % cb['one_thread']['value'] if cb.get('one_thread') else '')
    if 'collision_stage' in b:
        print(ind, '   collision_stage', b['collision_stage'])
    r = b.get('roofline')
    if r:
        bd = r.get('binding') or {}
        print(ind, '   roofline frac %.4f' % r['frac'], 'binding: %.0f GB/s of %.0f (%.3f)' % (bd['traffic_model_gbs'], bd['l2_peak_gbs_measured'], bd['frac_of_l2_peak']) if bd else '',
              (bd.get('per_ray') or ''))
    if c.get('exchange') and not c['exchange'].startswith('none'):
        print(ind, '   exchange:', c['exchange'][:110], '| timed out:', c.get('exchange_timed_out'))


for path in sys.argv[1:]:
    txt = [l for l in open(path).read().splitlines() if l.startswith('{')]
    d = json.loads(txt[-1])
    print(path, 'n_gpus', d.get('n_gpus'), 'steps', d.get('steps'), 'clocks', d.get('clocks'))
    show(d)
    for b in d.get('secondary', []):
        show(b, '  ')
