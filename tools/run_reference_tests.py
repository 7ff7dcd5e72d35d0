"""Run the reference's OWN test-suite, unmodified, against the drop-in modules.

The reference's tests import ``deepfake_detection``, ``frame_analysis``, ``model``, ``backend_server`` and
``face_detection`` by those bare names; this runner aliases the first four to ``dfd_b200.<name>`` in ``sys.modules`` and
leaves ``face_detection`` unresolved (face detection is out of scope, north_star: boxes are inputs), then runs pytest on
the test files and writes a pass / fail table.

The test FILES are not part of this repository (reference sources are never copied into history):

    python tools/run_reference_tests.py --stage     # build container: copy /root/reference/tests/test_*.py into
                                                    # oracle/_ref/reftests/ (git-ignored, travels to the GPU box)
    python tools/run_reference_tests.py --out gpurun_out/reftests_r02.md      # GPU box: run them

This is a checker (test infrastructure), never on the product path.
"""
import argparse
import importlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "oracle", "_ref", "reftests")
REF_TESTS = "/root/reference/tests"
ALIASES = ("deepfake_detection", "frame_analysis", "model", "backend_server")

# why a reference test cannot pass against a replacement of the hot path only (checked by name, reported beside the outcome)
OUT_OF_SCOPE = {
    "face_detection": "imports face_detection (OpenCV SSD / Haar detector): out of scope, boxes are inputs (north_star)",
    "test_model_file_exists": "weights/best_model.pth is not shipped with the reference either",
    "test_weights_load": "weights/best_model.pth is not shipped with the reference either",
    "compute_frequency_features": "model.compute_frequency_features is a training-side helper off the hot path",
}


def stage():
    if not os.path.isdir(REF_TESTS):
        raise SystemExit(f"{REF_TESTS} is not present (GPU box?): stage in the build container")
    os.makedirs(STAGE, exist_ok=True)
    n = 0
    for f in sorted(os.listdir(REF_TESTS)):
        if f.startswith("test_") and f.endswith(".py"):
            shutil.copyfile(os.path.join(REF_TESTS, f), os.path.join(STAGE, f))
            n += 1
    open(os.path.join(STAGE, "__init__.py"), "w").close()
    print(f"staged {n} reference test files in {STAGE}")


class _Collector:
    def __init__(self):
        self.rows = []

    def pytest_runtest_logreport(self, report):
        if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
            msg = ""
            if report.outcome != "passed":
                msg = str(report.longrepr).strip().splitlines()[-1][:200] if report.longrepr else ""
            self.rows.append((report.nodeid, report.outcome, msg, getattr(report, "duration", 0.0)))


def run(out_path, select):
    import pytest
    if not os.path.isdir(STAGE):
        raise SystemExit(f"{STAGE} missing: run with --stage in the build container first")
    sys.path.insert(0, ROOT)
    import dfd_b200  # noqa: F401
    for name in ALIASES:
        sys.modules[name] = importlib.import_module(f"dfd_b200.{name}")
    col = _Collector()
    args = [STAGE, "-q", "-p", "no:cacheprovider", "--tb=line", "-o", "addopts=",
            "--rootdir", STAGE]
    if select:
        args += ["-k", select]
    rc = pytest.main(args, plugins=[col])
    srcs = {}
    for f in os.listdir(STAGE):
        if f.endswith(".py"):
            srcs[f] = open(os.path.join(STAGE, f)).read()
    n_pass = sum(1 for r in col.rows if r[1] == "passed")
    n_skip = sum(1 for r in col.rows if r[1] == "skipped")
    lines = ["# Reference test-suite (unmodified) against the drop-in modules", "",
             f"`python tools/run_reference_tests.py` - sys.modules aliases {', '.join(ALIASES)} -> dfd_b200.*; pytest exit code {rc}.",
             "", f"**{n_pass} passed, {n_skip} skipped, {len(col.rows) - n_pass - n_skip} failed of {len(col.rows)}**", "",
             "| test | outcome | s | note |", "|---|---|---|---|"]
    for nodeid, outcome, msg, dur in col.rows:
        note = msg.replace("|", "/")
        short = nodeid.split("::")[-1]
        body = ""
        fname = nodeid.split("::")[0].split("/")[-1]
        if outcome != "passed" and fname in srcs:
            # the test's source text, to classify by what it imports
            s = srcs[fname]
            i = s.find(f"def {short}(")
            j = s.find("\n    def ", i + 1)
            body = s[i:j if j > 0 else len(s)]
        for key, why in OUT_OF_SCOPE.items():
            if outcome != "passed" and (key in short or key in body):
                note = f"OUT OF SCOPE: {why}. {note}"
                break
        lines.append(f"| {nodeid.replace(STAGE + '/', '')} | {outcome} | {dur:.2f} | {note} |")
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:5]))
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reftests_r02.md"))
    ap.add_argument("-k", default="")
    a = ap.parse_args()
    if a.stage:
        stage()
    else:
        sys.exit(run(a.out, a.k))
