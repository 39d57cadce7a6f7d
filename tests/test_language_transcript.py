"""The scene language's arithmetic, control flow, tuples, user functions and math builtins against the reference's own
interpreter (CPU). The script below prints a value per line; the reference's transcript (oracle/_ref/ref_render prints while
it parses) is committed as tests/golden/language_transcript.txt -- regenerate with
`python tests/test_language_transcript.py --make-golden` where /root/reference is compiled. The host interpreter must print
the same lines, character for character (the math builtins work on `float` like builtin_math.cpp:15-84, print() uses the
same 6-significant-digit format), up to the first error line, which carries the same message."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_util as ou  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "language_transcript.txt")

SCRIPT = """a = 3; b = 4.5;
print(a + b); print(a - b); print(a * b); print(a / b); print(a / 2); print(7 % 3); print(-a);
print(min(a, b)); print(clamp(5.5, 0, 1)); print(clamp(-2, -1, 1)); print(sqrt(2.0)); print(pow(2, 0.5));
print(sin(0.3)); print(cos(0.3)); print(tan(0.3)); print(asin(0.3)); print(acos(0.3)); print(atan(0.3)); print(atan(1.5, -0.5));
print(sin(100.0)); print(cos(3.1415926536)); print(pow(10, -3)); print(sqrt(1e-10)); print(1 / 3); print(2.0 / 3.0); print(1e6 * 1e3);
v = Vector(1, 2, 3); w = Vector(-2, 0.5, 4); p = Point(0.5, -1, 2);
print(dot(v, w)); c = cross(v, w); print(getX(c)); print(getY(c)); print(getZ(c));
print(distance(p, Point(1, 1, 1)));
t = (1, 2.5, "x": 7, (8, 9));
print(numElements(t)); print(t[0]); print(t[1]); print(t["x"]); print(t[2][1]);
addItem(t, 42); print(numElements(t)); print(t[3]);
i = 0; acc = 0;
for (i = 0; i < 5; ++i) { acc = acc + i * i; }
print(acc);
if (acc > 20) print(1); else print(0);
function f(x, y) { return x * y + 1; }
print(f(3, 4)); print(f("y": 2, "x": 5));
k = 10; k += 5; print(k); k -= 3; print(k); k *= 2; print(k); k /= 4; print(k);
print(1 < 2); print(2 <= 2); print(3 == 3); print(3 != 4); print(1 > 2 || 2 > 1); print(1 > 2 && 2 > 1); print(!(1 > 2));
print(random()); print(random());
print(undefinedName);
"""


def reference_transcript(d):
    with open(os.path.join(d, "lang.txt"), "w") as f:
        f.write(SCRIPT)
    return subprocess.run([os.path.join(ou.REF_DIR, "ref_render"), "lang.txt", "out.bin", "1", "8", "8"], cwd=d,
                          capture_output=True, text=True).stdout


def host_transcript(d):
    path = os.path.join(d, "lang.txt")
    with open(path, "w") as f:
        f.write(SCRIPT)
    code = ("import sys; sys.path.insert(0, %r)\nfrom slr_b200 import capi\ntry:\n    capi.read_scene(%r)\n"
            "except capi.SlrError as e:\n    print('ERROR:', str(e))\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), path))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout


def compare(got, want):
    got = [l.strip() for l in got.splitlines() if l.strip()]
    want = [l.strip() for l in want.splitlines() if l.strip()]
    assert len(got) == len(want) and len(want) > 50, (len(got), len(want))
    assert got[:-1] == want[:-1]
    # the last line is the error both interpreters stop at: same message (the host adds the file and line)
    assert got[-1].startswith("ERROR:") and want[-1].rstrip(".") in got[-1], (got[-1], want[-1])


def test_language_transcript_matches_golden(tmp_path):
    compare(host_transcript(str(tmp_path)), open(GOLDEN).read())


@pytest.mark.skipif(not ou.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_language_transcript_matches_live_reference(tmp_path):
    compare(host_transcript(str(tmp_path)), reference_transcript(str(tmp_path)))


if __name__ == "__main__" and "--make-golden" in sys.argv:
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        text = reference_transcript(d)
    with open(GOLDEN, "w") as f:
        f.write(text)
    print(len(text.splitlines()), "lines written to", GOLDEN)
