# quick GPU check of the update kernel: parity statistics (humanoid + dog) and the N=1 bench
timeout 300 python tools/parity_stats.py args/run_humanoid3d_spinkick_args.txt 200 2>&1 | tail -6
timeout 300 python tools/parity_stats.py args/train_dog3d_trot_args.txt 120 2>&1 | tail -6
timeout 300 python bench.py --steps 40 --warmup 4 2>&1 | tail -1 | cut -c1-400
