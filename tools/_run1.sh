timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for b in 1 2 4 8; do echo "resnet stem bands $b"; TLXCV_DEBUG_STEM_BANDS=$b timeout 120 python tools/stem_time.py resnet 2>&1 | tail -1; done
for b in 1 2 4 8; do echo "mobilenet stem bands $b"; TLXCV_DEBUG_STEM_BANDS=$b timeout 120 python tools/stem_time.py mobilenet 2>&1 | tail -1; done
for b in 2 3 5 7 10; do echo "darknet stem bands $b"; TLXCV_DEBUG_STEM_BANDS=$b timeout 120 python tools/stem_time.py darknet 2>&1 | tail -1; done
echo default; timeout 120 python tools/stem_time.py darknet 2>&1 | tail -1
timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 2>&1 | head -1
timeout 200 python tools/quick_prof.py --model resnext50_32x4d --batch 256 --size 224 2>&1 | head -1
