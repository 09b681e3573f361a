timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ab in 0 32 0 32; do TLXCV_DEBUG_ABLATE_STEM=$ab timeout 120 python tools/stem_time.py darknet 2>&1 | tail -1 | cut -c1-120; done
for ab in 0 32 0 32; do TLXCV_DEBUG_ABLATE_STEM=$ab timeout 120 python tools/stem_time.py mobilenet 2>&1 | tail -1 | cut -c1-120; done
