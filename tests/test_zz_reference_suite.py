"""The reference's OWN test-suite, unmodified, run against machupx_b200 ("run the 61 reference tests
against the new package").

`oracle/stage_ref.py` stages /root/reference/test byte for byte into oracle/_ref/test (git-ignored; it travels to the GPU box with
the snapshot like liblls.so; /root/reference itself is never read here).  The suite is copied into a scratch directory, a
two-line conftest resolves `import machupX` to this package, and pytest runs there in its own process:

* on CPU (`-m "not gpu"`) the device objects are the oracle stand-ins of tests/oracle_backend.py, so what is proven is the HOST
  mirror: input format, geometry, atmosphere / wind getters, section getters, drivers, result dictionaries;
* on the GPU box (`-m gpu`) nothing is replaced: every solve of the reference's tests goes through liblls.so.

Expected: 61 of the 62 non-CLI tests pass -- exactly the tests the unmodified reference itself passes on the airfoil_db stand-in
(SURVEY.md section 8c).  The 62nd, `test_swept_wing_with_aerodynamic_twist`, solves a wing of "database" airfoils whose golden
numbers depend on how the Qhull build of the golden's generation run broke the ties of the tables' exactly rectangular cells; it
is deselected here and restated, with those diagonals pinned, in tests/test_database_airfoils.py (where it passes to 1e-10 on
the oracle and on the device).  Run as it stands, with this image's scipy, it misses its golden by 3.5e-4 -- which the last test
below records, so that the claim stays checked.  `test/main_tests` drives the command-line front end (out of scope)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref", "test")
DATABASE_TEST = "test/scene_tests/test_solve_forces.py::test_swept_wing_with_aerodynamic_twist"

CONFTEST = '''import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import machupx_b200
sys.modules["machupX"] = machupx_b200          # the reference's tests say `import machupX as MX`
if os.environ.get("LLS_REFSUITE_DEVICE") == "oracle":
    import machupx_b200.device as device
    import oracle_backend
    device.DeviceScene = oracle_backend.OracleDeviceScene
    device.BatchDeviceScene = oracle_backend.OracleBatchDeviceScene
'''


def _run_suite(tmp_path, device, extra=()):
    if not os.path.isfile(os.path.join(STAGED, "input_for_testing.json")):
        pytest.skip("reference test-suite not staged (oracle/stage_ref.py needs /root/reference)")
    work = tmp_path / "refsuite"
    shutil.copytree(STAGED, work / "test", ignore=shutil.ignore_patterns("__pycache__", ".pytest_cache"))
    (work / "conftest.py").write_text(CONFTEST.format(root=ROOT))
    env = dict(os.environ, LLS_REFSUITE_DEVICE=device, PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "test", "--deselect", "test/main_tests", "-q", "--no-header", "-p", "no:cacheprovider",
           "-rf", *(extra or ("--deselect", DATABASE_TEST))]
    proc = subprocess.run(cmd, cwd=str(work), env=env, capture_output=True, text=True, timeout=1500)
    out = proc.stdout + proc.stderr
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|error|errors|skipped)", out.splitlines()[-1] if out.strip() else "")}
    failed = re.findall(r"^FAILED (\S+)", out, flags=re.M)
    return counts, failed, out


def _check(counts, failed, out):
    assert failed == [], "unexpected failures of the reference's own tests:\n" + out[-6000:]
    assert counts.get("passed") == 61 and not counts.get("failed") and not counts.get("error") and not counts.get("errors"), out[-3000:]


def test_reference_suite_against_the_host_mirror(tmp_path):
    counts, failed, out = _run_suite(tmp_path, "oracle")
    _check(counts, failed, out)


def test_the_database_test_as_it_stands_misses_its_golden_by_the_triangulation_only(tmp_path):
    """Unpinned, the reference's database test runs through (database airfoils are implemented) and lands within 1e-3 of its
    golden lift: the distance is the diagonal choice of three table cells (tests/test_database_airfoils.py)."""
    counts, failed, out = _run_suite(tmp_path, "oracle", extra=("-k", "test_swept_wing_with_aerodynamic_twist", "-s"))
    assert counts.get("failed", 0) + counts.get("passed", 0) == 1 and failed in ([DATABASE_TEST], []), out[-3000:]   # another Qhull may hit the golden
    m = re.search(r'"FL": (-?[0-9.]+)', out)
    assert m is not None and abs(float(m.group(1)) + 1.921431422738798) < 2e-3, out[-3000:]


@pytest.mark.gpu
def test_reference_suite_against_the_cuda_path(tmp_path):
    counts, failed, out = _run_suite(tmp_path, "cuda")
    _check(counts, failed, out)
    rec = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(rec):
        with open(os.path.join(rec, "reference_suite_gpu.txt"), "w") as fh:
            fh.write(out)
