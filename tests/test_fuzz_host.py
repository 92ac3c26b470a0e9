"""A short run of tools/fuzz_host.py in the CPU tier: corrupted containers, parameter sets, slice headers, payloads,
truncated / dropped / swapped samples and the whole Decoder.retrieve path on damaged files must end in an error or in
frames -- never in a crash or a hang (the long campaigns under ASan + UBSan: profiles/r2_fuzz_asan_summary.txt)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_host_short_run(emu):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_host.py'), '250', '5'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'fuzz ok: 250 iterations' in r.stdout


def test_param_sweep_short_run(emu):
    """Random combinations of the generator's options: generator reconstruction == libavcodec == decode core
    (tools/param_sweep.py; the long runs are recorded in DESIGN.md 6)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'param_sweep.py'), '40', '77'], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert 'param sweep: 40 combinations, 0 mismatches' in r.stdout
