ABLATIONS=0,128,26,154,2,130 timeout 300 python tools/ablate_sweep.py l1conv3 l2conv3 l3conv3 l4conv3 2>/dev/null
ABLATIONS=0,128,10,138 timeout 300 python tools/ablate_sweep.py l2conv1 l3conv1 l2conv2 2>/dev/null
TLXCV_DEBUG_ABLATE=128 timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 2>&1 | head -1
timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 2>&1 | head -1
