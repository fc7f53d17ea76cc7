"""CPU: a short run of scripts/fuzz_emulator.py (random scenes under every flag combination through the CPU emulation
of the kernel phases against the compiled reference core): z-buffer bit-exact, image and gradients inside the stated
fp32 tolerances.  The long campaigns of the round (~400 000 scenes) are run by hand with other seeds."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_short_fuzz_campaign(ref_oracle):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_emulator.py"), "12345", "20"],
                         capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    summary = out.stdout.strip().splitlines()[-1]
    assert " 0 failures" in summary, out.stdout[-3000:]
    assert int(summary.split()[0]) > 500  # it did run a meaningful number of scenes
