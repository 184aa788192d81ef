"""The reference's OWN test executables, unmodified, linked against the drop-in library.

scripts/link_reference_tests.sh (build container) compiles test programs of the reference where they lie under
/root/reference/test -- fnft__poly_fmult*, fnft__poly_fmult2x2*, fnft__poly_chirpz, fnft__poly_eval,
fnft__poly_roots_fasteigen, the 19 fnft__akns_fscatter_test_<scheme> programs, fnft__nse_scatter_bound_states_test_bo,
fnft_version_test -- against include/ of the reference and links them with -lfnft, which resolves to
fnft_b200/lib/libfnft.so -> libfnft_b200.so (SONAME libfnft.so.0, CMakeLists.txt:170-199 of the reference).  The
binaries travel to the GPU box with the snapshot; here every one of them must exit 0, i.e. meet the bounds the
reference's authors wrote into it."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "reftests_bin")


def _programs():
    lst = os.path.join(BIN, "LINKED.txt")
    if not os.path.exists(lst):
        return []
    return [l.strip() for l in open(lst) if l.strip()]


@pytest.mark.parametrize("name", _programs() or ["<none built>"])
def test_unmodified_reference_program(name, tmp_path):
    if name == "<none built>":
        pytest.skip("tests/reftests_bin is empty: run scripts/link_reference_tests.sh in the build container")
    exe = os.path.join(BIN, name)
    env = dict(os.environ)
    libdir = os.path.join(os.path.dirname(HERE), "fnft_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "libfnft.so.0")):  # the SONAME link did not travel: recreate it
        libdir = str(tmp_path)
        os.symlink(os.path.join(os.path.dirname(HERE), "fnft_b200", "lib", "libfnft_b200.so"),
                   os.path.join(libdir, "libfnft.so.0"))
    env["LD_LIBRARY_PATH"] = libdir + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (name, r.stdout[-1500:], r.stderr[-1500:])
