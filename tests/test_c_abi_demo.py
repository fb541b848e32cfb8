"""The C ABI used from plain C (examples/c_abi_demo.c): compiles against include/susnet_b200.h with gcc on every box;
runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "c_abi_demo")


def _compile():
    from sus_net_b200 import build as B

    B.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["/usr/bin/gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", EXE, "-L", os.path.join(ROOT, "sus_net_b200"),
           "-lsusnet_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{os.path.join(ROOT, 'sus_net_b200')}",
           f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_c_demo_compiles_against_the_header():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_c_demo_runs_on_the_gpu():
    _compile()
    out = subprocess.run([EXE, "20000", "150"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout and "env-steps/s" in out.stdout
