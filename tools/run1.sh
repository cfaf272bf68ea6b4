set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
WLS=cfg2 bash profiles/r1_sweep_staged2.sh
bash profiles/r1_exp_scaling.sh
python tools/dbg_timeline.py 8
python tools/dbg_timeline.py 32
