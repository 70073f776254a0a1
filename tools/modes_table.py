"""Markdown table of a tools/bench_modes.py JSON dump: modes_table.py <modes.json>"""
import json
import sys

rows = json.load(open(sys.argv[1]))
print('| config | paths × steps | kernel s | api s | path-steps/s (kernel) | HBM GB/s '
      '(algorithmic bytes / kernel s) | of 6554 GB/s measured copy peak |')
print('|---|---|---|---|---|---|---|')
for r in rows:
    if 'path_steps_per_s_kernel' in r:
        hb = r.get('hbm_gbs_kernel')
        print('| %s | %.0e × %d | %.4f | %.4f | %.3e | %s | %s |' % (
            r['config'], r['paths'], r['steps'], r['seconds_kernel'], r['seconds_api'],
            r['path_steps_per_s_kernel'], '%.0f' % hb if hb and hb > 10 else '—',
            '%.2f' % r['hbm_frac_of_measured_copy_peak'] if hb and hb > 10 else '—'))
print()
print('| resident-slab summary | seconds | GB/s | of copy peak |')
print('|---|---|---|---|')
for r in rows:
    if 'GBps' in r:
        print('| %s | %.5f | %.0f | %.2f |' % (r['config'], r['seconds'], r['GBps'],
                                             r['frac_of_copy_peak']))
print()
print('| host-facing run | paths × steps | wall s | path-steps/s (wall) | D2H GB/s (whole call) |')
print('|---|---|---|---|---|')
for r in rows:
    if 'seconds_wall' in r:
        print('| %s | %.0e × %d | %.4f | %.3e | %.1f |' % (
            r['config'], r['paths'], r['steps'], r['seconds_wall'],
            r['path_steps_per_s_wall'], r['d2h_GBps']))
