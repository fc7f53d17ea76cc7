"""Generates tests/golden/dropin_reference.json: loss / energy trajectories of the REFERENCE's own examples, run here
with the reference's own Cython extension (built in a scratch copy, see below), through tests/dropin/runner.py.

    D=/tmp/dref; mkdir $D; cp -r /root/reference/{deodr,C++,setup.py,readme.md,tests} $D; chmod -R u+w $D
    (cd $D && python setup.py build_ext --inplace)
    DEODR_STAGED_REFERENCE=$D python tests/golden/make_dropin_golden.py

The GPU acceptance test (tests/test_dropin_reference.py) replays the same examples with deodr_b200 bound as
`deodr.differentiable_renderer_cython` and compares the trajectories."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
RUNNER = os.path.join(os.path.dirname(HERE), "dropin", "runner.py")


def run(*args):
    out = subprocess.run([sys.executable, RUNNER, "ref", *[str(a) for a in args]], capture_output=True, text=True, check=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def main():
    golden = {"soup": {}, "hand_depth": {}, "hand_rgb": {}}
    for clockwise in (0, 1):
        for antialiase_error in (0, 1):
            golden["soup"][f"cw{clockwise}_err{antialiase_error}"] = run("soup", clockwise, antialiase_error, 50)
    for lib in ("none", "pytorch"):
        golden["hand_depth"][lib] = run("hand_depth", lib, 50)
    golden["hand_rgb"]["none"] = run("hand_rgb", "none", 50)
    with open(os.path.join(HERE, "dropin_reference.json"), "w") as f:
        json.dump(golden, f)
    print({k: list(v) for k, v in golden.items()})


if __name__ == "__main__":
    main()
