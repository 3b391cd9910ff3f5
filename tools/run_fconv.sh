cd /root/repo; mkdir -p gpurun_out
REPS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:fconv --launch-skip 7 -c 4 -f -o gpurun_out/fconv_p python tools/fconv_prof.py > gpurun_out/fconv_ncu.log 2>&1; echo ncu=$?
