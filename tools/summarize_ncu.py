"""Summarise an ncu report (run here, no GPU needed): headline metrics, stall breakdown, instruction-cache counters.
  python tools/summarize_ncu.py gpurun_out/x.ncu-rep "workload description" [launch index] > profiles/x.json
"""
import csv, json, subprocess, sys

rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1   # which captured launch (row) to summarise
val = rows[2:][which]
m = {h: (v, u) for h, u, v in zip(hdr, units, val)}


def g(k):
    v, u = m.get(k, ('', ''))
    try:
        return float(v.replace(',', ''))
    except ValueError:
        return v


def scaled(k):
    """dram byte counters come back in a scaled unit (Mbyte / Gbyte): normalise to bytes."""
    v, u = m.get(k, ('', ''))
    f = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)
    return float(v.replace(',', '')) * f


out = {
    'kernel': m['Kernel Name'][0],
    'workload': workload,
    'grid': m['Grid Size'][0], 'block': m['Block Size'][0],
    'duration_ms_under_ncu': g('gpu__time_duration.sum') / ({'usecond': 1e3, 'msecond': 1, 'second': 1e-3, 'nsecond': 1e6}.get(m['gpu__time_duration.sum'][1], 1)),
    'registers_per_thread': g('launch__registers_per_thread'),
    'warp_instructions_executed': g('smsp__inst_executed.sum'),
    'dram_bytes_read': scaled('dram__bytes_read.sum'), 'dram_bytes_write': scaled('dram__bytes_write.sum'),
    'dram_throughput_pct_of_peak': g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    'sm_warps_active_pct_of_peak': g('sm__warps_active.avg.pct_of_peak_sustained_active'),
    'smsp_issue_active_pct': g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
    'warp_cycles_per_issued_instruction': g('smsp__average_warp_latency_per_inst_issued.ratio'),
    'l1_hit_pct': g('l1tex__t_sector_hit_rate.pct'), 'l2_hit_pct': g('lts__t_sector_hit_rate.pct'),
    'sm_icache_hit_pct (sm__icc_request_hit_rate)': g('sm__icc_request_hit_rate.pct'),
    'sm_icache_requests': g('sm__icc_requests.sum'),
    'gpc_icache_instruction_requests (gcc)': g('gcc__cache_requests_type_instruction.sum'),
    'gpc_icache_instruction_requests_pct_of_peak': g('gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed'),
    'stall_cycles_per_issued_instruction': {
        k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(g(k), 3)
        for k in hdr if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and isinstance(g(k), float) and g(k) >= 0.005},
}
out['dram_bytes_per_launch'] = out['dram_bytes_read'] + out['dram_bytes_write']
print(json.dumps(out, indent=1))
