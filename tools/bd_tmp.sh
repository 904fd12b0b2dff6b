o=gpurun_out
python tools/step_breakdown.py --model sd-turbo --batch 64 --mode static --out $o/r2H_sd64s.json > $o/r2H_sd64s.txt 2>&1
python tools/crit_path.py $o/r2H_sd64s.json 45
python tools/step_breakdown.py --batch 8 --mode static --out $o/r2H_b8s.json > $o/r2H_b8s.txt 2>&1
python tools/crit_path.py $o/r2H_b8s.json 30
