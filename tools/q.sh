timeout 150 python tools/prof_episode.py --sweep --episodes 1048576 2>&1 | tail -2 | cut -c1-400
timeout 150 python tools/prof_episode.py --sweep --episodes 1048576 --launch-ticks 500 --compact-every 500 2>&1 | tail -2 | cut -c1-120
timeout 200 python -m pytest tests/test_init.py -m gpu -x -q 2>&1 | tail -2
