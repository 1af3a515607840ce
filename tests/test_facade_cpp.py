"""The C++ host facades (gf-pl-slam_b200/host: StVO::StereoFrame, StVO::StereoFrameHandler, PLSLAM::MapHandler)
compile against the C ABI, fail loudly without a device, and on a GPU reproduce the oracle when driven like the
reference's main loop (tests/cpp/facade_parity.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gf-pl-slam_b200", "host")
LIB = os.path.join(ROOT, "gf-pl-slam_b200", "lib")
ORC = os.path.join(ROOT, "oracle", "_build")
EXE = os.path.join(ROOT, "tests", "cpp", "_facade_parity")


def build_exe():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)
    src = [os.path.join(ROOT, "tests", "cpp", "facade_parity.cpp")] + [os.path.join(HOST, f) for f in
                                                                       ("stereoFrame.cpp", "stereoFrameHandler.cpp", "mapHandler.cpp")]
    newest = max(os.path.getmtime(s) for s in src + [os.path.join(LIB, "libgfpl.so")])
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest:
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", EXE] + src +
                              ["-L", LIB, "-lgfpl", "-L", ORC, "-lgfpl_oracle", f"-Wl,-rpath,{LIB}", f"-Wl,-rpath,{ORC}", "-pthread"])
    return EXE


def test_facade_compiles_and_fails_loudly_without_gpu():
    exe = build_exe()
    out = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_facade_parity_on_gpu():
    exe = build_exe()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0 and "FACADE PARITY OK" in out.stdout, out.stdout + out.stderr
