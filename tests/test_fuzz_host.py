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
