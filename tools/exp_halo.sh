run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_quick_bench.py --steps 4 --warmup 2 --table 2>&1 | grep -E "best|up_3.conv|up_3.ace_1|up_2.conv|up_2.ace_1|up_1.conv_1|up_1.ace_1|conv_img" ; }
run A=0
run CHB_WSTAT_MINHALO=4
run CHB_WSTAT_MINHALO=4 CHB_NHALO64=4
run CHB_WSTAT_MINHALO=7 CHB_NHALO64=4
run CHB_NHALO128=3
run CHB_NHALO128=4
run A=0
