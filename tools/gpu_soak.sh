set -x
timeout 600 python tools/soak.py args/train_humanoid3d_spinkick_args.txt 6000 2>&1 | tail -2
timeout 600 python tools/soak.py args/train_dog3d_trot_args.txt 3000 2>&1 | tail -2
timeout 600 python tools/soak.py args/train_humanoid3d_walk_args.txt 3000 2>&1 | tail -2
