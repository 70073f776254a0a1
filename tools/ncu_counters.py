"""Counters of one kernel from an `ncu --page raw --csv` dump -> JSON for bench.py.

    python tools/ncu_counters.py <raw.csv> <paths> <steps> <commit> "<command>" > profiles/rNN_<kernel>_counters.json

Everything per path-step is (counter of the launch) / (paths x steps [/ 32 for
warp instructions]); nothing here is typed in by hand.
"""
import csv
import json
import sys


def main():
    raw, paths, steps, commit, command = sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), sys.argv[4], sys.argv[5]
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def f(name):
        return float(m[name][0].replace(',', ''))

    def mb(name):
        v, u = m[name]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        return float(v.replace(',', ''))*scale

    warp_steps = paths*steps/32
    inst = f('smsp__inst_executed.sum')
    cyc_active = f('sm__cycles_active.sum')
    fp64_pct = f('sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active')
    # FP64 pipe: 16 lanes per sub-partition = 2 warp instructions per SM cycle at peak
    fp64_inst = fp64_pct/100*2*cyc_active
    stalls = {h.split('issue_stalled_')[1].split('_per_issue')[0]: float(v[0])
              for h, v in m.items() if 'smsp__average_warps_issue_stalled_' in h
              and h.endswith('per_issue_active.ratio') and v[0] not in ('', 'n/a')}
    out = {
        'kernel': m['Kernel Name'][0], 'commit': commit, 'source': raw, 'command': command,
        'paths': paths, 'steps': steps,
        'duration_ms': f('gpu__time_duration.sum')*{'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[m['gpu__time_duration.sum'][1]],
        'sm_mhz': f('device__attribute_clock_rate')/1e3,
        'registers_per_thread': f('launch__registers_per_thread'),
        'warps_active_pct': f('sm__warps_active.avg.pct_of_peak_sustained_active'),
        'warp_instr_per_path_step': inst/warp_steps,
        'fp64_pipe_instr_per_path_step': fp64_inst/warp_steps,
        'cycles_per_warp_step': f('smsp__cycles_elapsed.sum')/warp_steps,
        'ncu_fp64_pipe_pct': f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
        'ncu_alu_pipe_pct': f('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'),
        'ncu_fmaheavy_pipe_pct': f('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'),
        'ncu_xu_pipe_pct': f('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
        'ncu_lsu_data_pipe_pct': f('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
        'ncu_issue_active_pct': f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'shared_bank_conflict_wavefronts': f('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
        'dram_bytes_read': mb('dram__bytes_read.sum'), 'dram_bytes_written': mb('dram__bytes_write.sum'),
        'dram_bytes_per_path': (mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'))/paths,
        'stalls_per_issue': dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8]),
    }
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
