mkdir -p gpurun_out
export BJXA_B200_STEREO=direct
A="tools/prof_decode.py --mix P0,P1 --streams 2048 --seconds 30 --bits 8 --ch 2 --steps 1 --warmup 1"
timeout 300 python $A > gpurun_out/ncu_stereo_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xa_decode -o gpurun_out/stereo_direct_r1 -f python $A > gpurun_out/ncu_stereo.log 2>&1
echo rc=$?
