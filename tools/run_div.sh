cd /root/repo/tools
echo default; timeout 200 python diverge_check.py 2>&1 | tail -3
echo serial; SEGB200_NO_GRAPH=1 SEGB200_WGRAD_STREAM=0 timeout 200 python diverge_check.py 2>&1 | tail -3
echo nograph; SEGB200_NO_GRAPH=1 timeout 200 python diverge_check.py 2>&1 | tail -3
echo eff70; SEGB200_TCONV_MIN_EFF=70 timeout 200 python diverge_check.py 2>&1 | tail -3
echo lr1e-4; LR=1e-4 timeout 200 python diverge_check.py 2>&1 | tail -3
