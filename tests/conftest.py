import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def kat_lines():
    # SURVEY.md App. D dictionary (prefix mode => also 甲甲甲:0, 戊:0; size 359)
    return "甲 100,甲甲 50,甲甲甲甲 7,乙 40,乙丙 30,乙丙丁 5,丙 20,丁 60,丙丁 25,戊己 9,己 3,庚 8,辛 2".split(",")


@pytest.fixture(scope="session")
def kat_emit():
    emit = {s: {c: -5.0 - i * 0.1 for i, c in enumerate("甲乙丙丁己庚辛")} for s in "BMES"}
    emit["S"]["壬"] = -6.0
    return emit


@pytest.fixture(scope="session")
def small_synth():
    from jieba_go_b200 import synth
    sd = synth.make_dictionary(n_words=6000, seed=synth.SEED_BASE + 11, total_freq=2.0e6, max_len=9)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 12)
    return sd, emit
