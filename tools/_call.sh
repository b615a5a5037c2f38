cd /root/repo
export B200TRK_SD_FILL=ldg
B200TRK_SD_TC=1 timeout 300 python -m pytest tests -m gpu -q -k "sd or dimp or hinge or l2 or track" 2>&1 | tail -4
cd tools
timeout 200 python sd_tc_check.py 2>&1 | grep "tc=1\|rel" | cut -c1-150
timeout 120 python sd_tc_dbg.py 2>&1 | grep "dbg 0" | head -2
