#!/usr/bin/env python
"""After a GPU visit: turn the pending (xfail) GPU tests that PASSED on the B200 into plain tests.

    python tools/promote_pending.py gpurun_out/junit_*.xml          # dry run: prints what would change
    python tools/promote_pending.py --apply gpurun_out/junit_*.xml

Input: the junit files tools/gpu_train_checks.sh writes (pytest --runxfail --junitxml: real pass / fail per case).  A test
function is promoted when EVERY one of its parametrised cases passed: its `@PENDING` decorator is removed; a file whose
module-level `pytestmark` carries the xfail marker is promoted as a whole (marker dropped) when every case of the file passed.
Anything that failed, errored or did not run keeps its marker, and is listed."""
import argparse
import collections
import os
import re
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def outcomes(paths):
    """{test file: {function: [ok, ok, ...]}} from junit testcases (classname tests.test_x, name func[params])."""
    res = collections.defaultdict(lambda: collections.defaultdict(list))
    for p in paths:
        for tc in ET.parse(p).getroot().iter("testcase"):
            mod = tc.get("classname", "").split(".")[-1]
            fn = tc.get("name", "").split("[")[0]
            bad = any(ch.tag in ("failure", "error", "skipped") for ch in tc)
            res[os.path.join("tests", mod + ".py")][fn].append(not bad)
    return res


def promote(path, funcs, apply):
    src = open(os.path.join(ROOT, path)).read()
    passed = sorted(f for f, oks in funcs.items() if oks and all(oks))
    failed = sorted(f for f, oks in funcs.items() if not (oks and all(oks)))
    defined = set(re.findall(r"^def (test_\w+)", src, flags=re.M))
    not_run = sorted(defined - set(funcs))
    new = src
    module_mark = re.search(r"^pytestmark = \[pytest\.mark\.gpu,\n\s+pytest\.mark\.xfail\([^\n]*\)\]\n", src, flags=re.M)
    if module_mark:
        if not failed and not not_run:
            new = new.replace(module_mark.group(0), "pytestmark = pytest.mark.gpu\n")
            what = "whole file promoted"
        else:
            what = "file keeps its marker (module-level xfail; failed or missing cases)"
    else:
        n = 0
        for f in passed:
            new, k = re.subn(r"^@PENDING\n((?:@pytest\.mark\.parametrize\([^\n]*\n(?:\s+[^\n]*\n)*?)*)(def %s\()" % re.escape(f), r"\1\2", new, flags=re.M)
            n += k
        what = f"{n} test functions promoted"
    print(f"{path}: {what}; passed={len(passed)} failed={failed} not_run={not_run}")
    if apply and new != src:
        open(os.path.join(ROOT, path), "w").write(new)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--apply", action="store_true")
    ap.add_argument("junit", nargs="+")
    a = ap.parse_args()
    for path, funcs in sorted(outcomes(a.junit).items()):
        if os.path.exists(os.path.join(ROOT, path)):
            promote(path, funcs, a.apply)
    return 0


if __name__ == "__main__":
    sys.exit(main())
