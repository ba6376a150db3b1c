"""REPL transcript parity: one command file through the reference's own REPL (src/main.cpp:40-76, built
unmodified into oracle/_ref/pixiu_repl) and through the same REPL written against include/PiXiuCtrl.hpp
(tests/cpp/facade_repl.cpp): identical stdout, including the saved-bytes lines (main.cpp:67-70)."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "pixiu_repl")
OURS = os.path.join(ROOT, "tests", "cpp", "facade_repl")


def transcript(seed=7, n=400):
    """SET / GET lines in the style of README.md:70-96: URL values sharing long substrings, repeated keys, misses"""
    rng = random.Random(seed)
    hosts = ["www.zhihu.com/question", "news.qq.com/a/20161101", "sports.qq.com/a/20161102", "github.com/Thunderchen/PiXiu/issues"]
    keys, lines = [], []
    for i in range(n):
        r = rng.random()
        if r < 0.65 or not keys:
            k = "K%d" % rng.randrange(10 ** rng.randint(1, 5)) if rng.random() < 0.8 else rng.choice(keys)
            v = "https://%s/%d" % (rng.choice(hosts), rng.randrange(10 ** 8))
            if rng.random() < 0.2:
                v += "?ref=" + "ab" * rng.randint(1, 40)          # self-overlapping runs
            if rng.random() < 0.1:
                v += "&pad=" + "x" * rng.choice([250, 251, 252, 255, 256, 300])   # run lengths around the 251 / 255 boundaries
            keys.append(k)
            lines.append("SET %s::%s" % (k, v))
        elif r < 0.9:
            lines.append("GET %s" % rng.choice(keys))
        else:
            lines.append("GET missing%d" % i)
    lines.append("~")
    return "\n".join(lines) + "\n"


def test_reference_repl_runs_and_saves_bytes():
    """(CPU) the reference binary itself: README.md:70-96 style session"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/pixiu_repl not built (no /root/reference here)")
    out = subprocess.run([REF], input="SET BOBO::https://www.zhihu.com/question/55439090\n"
                                       "SET BOBO1::https://www.zhihu.com/question/22454692\nGET BOBO1\n~\n",
                         capture_output=True, text=True, timeout=60).stdout
    assert "BOBO1::https://www.zhihu.com/question/22454692" in out and " 29" in out


@pytest.mark.gpu
def test_repl_transcript_matches_reference():
    assert os.path.exists(REF), "oracle/_ref/pixiu_repl missing: run __graft_entry__.build() where /root/reference is mounted"
    assert os.path.exists(OURS), "run __graft_entry__.build() first"
    text = transcript()
    a = subprocess.run([REF], input=text, capture_output=True, text=True, timeout=300)
    b = subprocess.run([OURS], input=text, capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-500:], b.stderr[-500:])
    assert a.stdout.count("Command: ") == text.count("\n")
    if a.stdout != b.stdout:
        la, lb = a.stdout.split("Command: "), b.stdout.split("Command: ")
        for i, (x, y) in enumerate(zip(la, lb)):
            assert x == y, f"command {i} ({text.splitlines()[i - 1]!r}): reference {x!r} vs ours {y!r}"
    assert a.stdout == b.stdout
