"""Random-intruder MCTS env leg alone: python tools/mctsrnd_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import bench  # noqa: E402

print(json.dumps(bench.bench_mctsrnd(0)))
