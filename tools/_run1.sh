timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in prev new prev new; do
echo "== $v"
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so ABLATIONS=0 timeout 300 python tools/ablate_sweep.py l1conv3 l2conv3 l3conv3 l4conv3 l3conv1 l2conv1 2>/dev/null
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 2>&1 | head -1
done
