"""The reference's Hector headers are compiled against oracle/shim/Eigen/mini_eigen.h (Eigen itself is absent from the
image).  tests/eigen_shim_check.cpp pins the stand-in's primitives to hand-written IEEE-754 float32 operation orders: the
operand order of Translation * Rotation applied to a point, AlignedScaling * Translation and its Affine inverse,
float -> int truncation of negatives, three-term row sums as p0 + (p1 + p2), the cofactor 3x3 inverse."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shim_primitives_match_hand_derived_float32(tmp_path):
    exe = str(tmp_path / "eigen_shim_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle", "shim"),
                           os.path.join(HERE, "eigen_shim_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert json.loads(out.stdout.strip().splitlines()[-1])["failures"] == 0
