"""CPU: stress of the host path's copy-thread crew (deodr_b200/csrc/host_crew.h): hundreds of batches back to back with
random chunk kinds, sizes, task granularities and widths, in the three coordination patterns of the host entry points
(plain, upload = chunks consumed in order as they complete, download = chunks gated one by one), every output checked.
A lost wake-up or a gate race would show up as a hang (pytest timeout) or a mismatch."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-C", os.path.join(HERE, "emul"), "libcrew.so"], check=True, capture_output=True)
    so = C.CDLL(os.path.join(HERE, "emul", "libcrew.so"))
    so.hook_crew_stress.restype = C.c_long
    so.hook_crew_stress.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    return so


@pytest.mark.timeout(120)
@pytest.mark.parametrize("workers", [1, 3, 7, 15])
def test_crew_stress(lib, workers):
    assert lib.hook_crew_stress(workers, 400, 1000 + workers, 64) == 0


@pytest.mark.timeout(120)
def test_crew_many_small_batches(lib):
    """Tiny batches back to back: the publish / sleep / wake protocol rather than the copy loops."""
    assert lib.hook_crew_stress(15, 4000, 7, 1) == 0


@pytest.mark.timeout(60)
def test_mirror_comparison_reports_a_difference(lib):
    assert lib.hook_crew_detects_difference(5) == 1
