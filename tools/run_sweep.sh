cd /root/repo
timeout 600 python tools/debug_handoff.py 2>&1 | grep -v "^   " | tail -30
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
