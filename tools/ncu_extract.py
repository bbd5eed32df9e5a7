"""Extract the headline metrics of every kernel in an .ncu-rep (ncu --set full) into a small json for profiles/.
usage: python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/out.json "<note>" """
import csv
import io
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']


def main(rep, out, note):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    ks = []
    for r in data:
        d = {'kernel': r[col['Kernel Name']][:120], 'grid': r[col['Grid Size']], 'block': r[col['Block Size']]}
        for k in KEEP:
            if k in col:
                d[k] = f'{r[col[k]]} {units[col[k]]}'.strip()
        ks.append(d)
    json.dump({'source': note, 'kernels': ks}, open(out, 'w'), indent=1)
    for d in ks:
        print(d['kernel'][:60], d.get('gpu__time_duration.sum'), 'dram r/w', d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum'))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
