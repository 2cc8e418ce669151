"""The C ABI driven from a plain-C host program (tests/c_host/c_abi_host.c): no Python objects, no torch types -- only
device pointers, sizes and a cudaStream_t cross the boundary, exactly what the reference's FFI layer would hand over.
The program checks itself against the C oracle and exits 0 when every case matched."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "c_host", "c_abi_host")


def test_c_host_program_is_built_and_links():
    """CPU: build() compiled it and the dynamic loader resolves the product library and the checker."""
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.join(HERE, "c_host")], check=True, capture_output=True)
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True, check=True).stdout
    assert "libbrainevent_b200.so" in out and "not found" not in out.split("libbrainevent_b200.so")[1].splitlines()[0]
    assert "libbe_oracle.so" in out


@pytest.mark.gpu
def test_c_host_program_matches_oracle():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all cases match the oracle" in r.stdout
    assert r.stdout.count(" ok") >= 8
