#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_lightgen.py tests/test_gpu_dropin.py -q 2>&1 | tail -3
python tools/ncu_frame.py 2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file /tmp/l.csv python tools/ncu_frame.py 2 > /dev/null 2>&1; grep -E "k_light" /tmp/l.csv | cut -d, -f5,15 | head
python tools/e2e_probe2.py gen 2>&1 | tail -2
