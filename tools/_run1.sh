timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in packed new packed new; do
echo "== $v"
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so ABLATIONS=0 timeout 300 python tools/ablate_sweep.py 2>/dev/null
done
for v in packed new packed new; do
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 2>&1 | head -1
done
for v in packed new; do
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model mobilenet_v2 --batch 512 --size 224 2>&1 | head -1
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model darknet53_det --batch 64 --size 608 2>&1 | head -1
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnext50_32x4d --batch 256 --size 224 2>&1 | head -1
done
