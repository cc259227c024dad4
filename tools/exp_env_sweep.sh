for w in 3840 7680 15360; do echo "window=$w"; MFHN_CATEGORIZE_WINDOW=$w python tools/exp_runs.py --degree 4 --occ 5 10:6 2>&1 | grep 'kernel": "runs' | cut -c1-170; done
echo "maxpad14"; MFHN_RUNS_MAXPAD=14 python tools/exp_runs.py --degree 4 --occ 5 10:6 2>&1 | grep 'kernel": "runs' | cut -c1-170
echo "refine4"; MFHN_RUNS_REFINE=4 python tools/exp_runs.py --degree 4 --occ 5 10:6 2>&1 | grep 'kernel": "runs' | cut -c1-170
echo "k5"; python tools/exp_runs.py --degree 5 --occ 4,5 10:6 16:8 6:6 24:10 2>&1 | grep -v 'kernel": "plane' | cut -c1-170
