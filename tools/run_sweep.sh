cd /root/repo
timeout 600 python tools/sweep_plain.py 2>&1 | tail -12
