#!/bin/bash
# dev: A/B of two prebuilt libraries + timing table + bit comparison + gpu tests: tools/try.sh old.so new.so
tools/ab.sh -a "--log2n 20 --no-extras --no-e2e" $1 $2 $1 $2 2>&1 | tail -8
python tools/timing.py run 19
python tools/lib_equal.py ab/base.so $2 20000 2>&1 | tail -14
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
