"""bench.py's CPU-only pieces: the C1 record (examples/pitch_detection.rs as shipped, timed on the oracle) must work with and
without --steps (the driver passes --steps; a bare `python bench.py` does not), and the reference arm's argument defaults."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_c1_record_without_and_with_steps():
    for steps in (None, 3):
        rec = bench.c1_record(argparse.Namespace(steps=steps))
        assert rec["unit"] == "frames/s" and rec["value"] > 0 and rec["higher_is_better"] is True
        assert rec["steps"] == (20 if steps is None else 3)
        case = rec["cases"]["sine_150hz_2048"] if "sine_150hz_2048" in rec["cases"] else next(iter(rec["cases"].values()))
        assert isinstance(case, dict)
