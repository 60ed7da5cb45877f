#!/bin/bash
# usage: gpu_tests.sh [pytest -k expression]
if [ -n "$1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -40
else
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
fi
