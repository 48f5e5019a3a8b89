"""The C++ host side (include/traccc_b200/seeding.hpp): compiles everywhere; on the GPU box
the program re-runs the reference's own seeding unit tests through the adapter."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_adapter_compiles():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")], stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(HERE, "cpp", "test_seeding_adapter"))


@pytest.mark.gpu
def test_reference_unit_tests_through_adapter():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(HERE, "cpp", "test_seeding_adapter")], capture_output=True,
                         text=True, timeout=120)
    print(out.stdout, out.stderr)
    assert out.returncode == 0
    assert out.stdout.count("[ OK ]") == 6
