timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in prev new; do
echo "== $v"
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so ABLATIONS=0,42 timeout 300 python tools/ablate_sweep.py l2conv3 l3conv3 l3conv1 l4conv3 l4conv1 l3conv2 l3conv2_nopair l2conv2 l2conv2_pair l2conv1 l4conv2 l4conv2_pair 2>/dev/null
done
for v in prev new prev new; do
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 2>&1 | head -1
done
for v in prev new; do
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model mobilenet_v2 --batch 512 --size 224 2>&1 | head -1
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model darknet53_det --batch 64 --size 608 2>&1 | head -1
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnext50_32x4d --batch 256 --size 224 2>&1 | head -1
done
