timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in prev new prev new; do
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnet50 --batch 256 --size 224 --out gpurun_out/prof_r50_$v.json 2>&1 | grep -E "^resnet50|downsample|layer4" | cut -c1-150
done
for v in prev new; do
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model mobilenet_v2 --batch 512 --size 224 2>&1 | head -1
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model darknet53_det --batch 64 --size 608 2>&1 | head -1
TLXCV_B200_LIB=$PWD/tlxcv_b200/lib_$v.so timeout 200 python tools/quick_prof.py --model resnext50_32x4d --batch 256 --size 224 2>&1 | head -1
done
