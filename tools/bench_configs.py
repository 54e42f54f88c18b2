#!/usr/bin/env python3
"""Full-size runs of the BASELINE.json configurations that are not the default bench.py workload (the
implementations live in bench.py: run_config_c4 / run_config_c5; `python bench.py --config c4|c5` prints the same
numbers in the bench contract).
usage: bench_configs.py {c4|c5} [N]      -> one JSON line (device-timed with CUDA events)"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
import bench  # noqa: E402


def main():
    which = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    out = bench.run_config_c4(n) if which == "c4" else bench.run_config_c5(n)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
