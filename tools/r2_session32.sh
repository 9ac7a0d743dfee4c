#!/bin/bash
cd "$(dirname "$0")/.."
python tools/stage_probe.py oracle/_ref/blobs/bounding_boxes.frt 2>&1 | tail -1 | cut -c1-400
python tools/stage_probe.py oracle/_ref/blobs/sibenik_surrogate.frt 400 500 4 2>&1 | tail -1 | cut -c1-400
python tools/stage_probe.py oracle/_ref/blobs/teapot.frt 2>&1 | tail -1 | cut -c1-400
python tools/stage_probe.py oracle/_ref/blobs/reflect_refract.frt 2>&1 | tail -1 | cut -c1-400
