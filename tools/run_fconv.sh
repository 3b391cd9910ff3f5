cd /root/repo; mkdir -p gpurun_out
for d in 0 8 12 13 14 15 24 28 31; do echo "dbg=$d"; SEGB200_FCONV_DBG=$d python tools/fconv_prof.py 2>&1 | grep wgrad; done
