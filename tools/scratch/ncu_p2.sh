mkdir -p gpurun_out
A="tools/prof_decode.py --mix P2 --streams 2048 --seconds 30 --bits 8 --ch 1 --steps 1 --warmup 1"
timeout 300 python $A > gpurun_out/ncu_p2_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xa_decode -c 1 -o gpurun_out/mono_p2_r1 -f python $A > gpurun_out/ncu_p2.log 2>&1
echo rc=$?
