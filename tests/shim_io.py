"""Dump a ldso_b200.synth.Window in the flat binary format tests/cpp/shim_test.cc reads, and build that test."""
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_test")


def dump(win, path):
    with open(path, "wb") as f:
        f.write(np.array([win.w, win.h, win.levels, win.nF, win.nP, win.nR], np.int32).tobytes())
        f.write(np.asarray(win.K, np.float64).tobytes())
        for i in range(win.nF):
            f.write(np.ascontiguousarray(win.Rcw[i], np.float64).tobytes())
            f.write(np.ascontiguousarray(win.tcw[i], np.float64).tobytes())
            f.write(np.ascontiguousarray(win.state_zero[i], np.float64).tobytes())
            f.write(np.ascontiguousarray(win.state[i], np.float64).tobytes())
            f.write(np.float32(win.ab_exposure[i]).tobytes())
            f.write(np.int32(win.frame_id[i]).tobytes())
            for l in range(win.levels):
                f.write(np.ascontiguousarray(win.pyramids[i][l], np.float32).tobytes())
        for a, dt in ((win.pt_host, np.int32), (win.pt_u, np.float32), (win.pt_v, np.float32), (win.pt_idepth, np.float32),
                      (win.pt_idepth_zero, np.float32), (win.pt_color, np.float32), (win.pt_weights, np.float32),
                      (win.res_begin, np.int32), (win.res_target, np.int32)):
            f.write(np.ascontiguousarray(a, dt).tobytes())


def build():
    src = os.path.join(ROOT, "tests", "cpp", "shim_test.cc")
    # always rebuilt (a second of g++): mtimes are arbitrary after a checkout / snapshot, a stale binary must never be run
    libdir = os.path.join(ROOT, "ldso_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", src, "-o", EXE, "-L" + libdir, "-lldso_b200", "-Wl,-rpath," + libdir, "-pthread"])
    return EXE
