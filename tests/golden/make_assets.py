"""Packs the reference's DATA files that the tests / smoke / bench need (character, controller, motion,
terrain JSON and arg files -- inputs of the hot path, not source code) into tests/golden/assets.tar.gz.
Run in the build container where /root/reference exists:  python tests/golden/make_assets.py
The GPU box has no /root/reference; tests unpack this archive instead (deepmimic_b200/assets.py)."""
import os, tarfile

REF = "/root/reference"
FILES = [
    "data/characters/humanoid3d.txt", "data/characters/dog3d.txt",
    "data/controllers/humanoid3d_ctrl.txt", "data/controllers/humanoid3d_phase_rot_ctrl.txt", "data/controllers/humanoid3d_rot_ctrl.txt",
    "data/controllers/dog3d_ctrl.txt", "data/controllers/dog3d_phase_ctrl.txt", "data/controllers/dog3d_phase_rot_ctrl.txt", "data/controllers/dog3d_rot_ctrl.txt",
    "data/motions/humanoid3d_spinkick.txt", "data/motions/humanoid3d_walk.txt", "data/motions/humanoid3d_run.txt", "data/motions/humanoid3d_backflip.txt",
    "data/motions/dog3d_trot.txt", "data/motions/dog3d_pace.txt",
    "data/terrain/plane.txt",
    "args/run_humanoid3d_spinkick_args.txt", "args/train_humanoid3d_spinkick_args.txt", "args/run_humanoid3d_walk_args.txt", "args/train_humanoid3d_walk_args.txt",
    "args/train_humanoid3d_run_args.txt", "args/train_humanoid3d_backflip_args.txt",
    "args/run_dog3d_trot_args.txt", "args/train_dog3d_trot_args.txt", "args/train_dog3d_pace_args.txt",
    # AMP task scenes (target / heading); their 48 MB clip dataset is replaced in the tests by the authored mini dataset below
    "args/train_amp_target_humanoid3d_locomotion_args.txt", "args/train_amp_heading_humanoid3d_locomotion_args.txt",
    # heading + get-up: its 4-clip dataset is small enough to ship whole (run and walk are above); strike: args only, run on the mini dataset
    "args/train_amp_heading_getup_humanoid3d_locomotion_getup_args.txt", "data/datasets/humanoid3d_clips_locomotion_getup.txt",
    "data/motions/humanoid3d_getup_facedown.txt", "data/motions/humanoid3d_getup_faceup.txt",
    "args/train_amp_strike_humanoid3d_walk_punch_args.txt",
]

# Authored here (not reference data): a cClipsController dataset over clips that are already in the archive, in the reference's format
# (R/data/datasets/humanoid3d_clips_locomotion.txt).  Tests pass it with --motion_file, which wins over the arg file's entry.
MINI_DATASET = "data/datasets/test_clips_mini.txt"
MINI_DATASET_TEXT = """{
	"Motions":
	[
		{"Weight": 20, "File": "data/motions/humanoid3d_run.txt"},
		{"Weight": 3, "File": "data/motions/humanoid3d_walk.txt"},
		{"Weight": 1, "File": "data/motions/humanoid3d_spinkick.txt"},
		{"File": "data/motions/humanoid3d_backflip.txt"}
	]
}
"""

# Authored as well: a 56-entry dataset with the SHAPE of R/data/datasets/humanoid3d_clips_locomotion.txt (56 clips, 48 MB -- too large to
# ship) over the two locomotion clips of the archive, with varying weights: BASELINE.json config 5 ("AMP target humanoid3d_locomotion, 4096
# envs") runs on it in bench.py and in the GPU tests (synthetic data of the reference's shape; every entry is loaded as a clip of its own).
SYN56_DATASET = "data/datasets/synthetic_locomotion_56.txt"
SYN56_DATASET_TEXT = "{\n\t\"Motions\":\n\t[\n" + ",\n".join(
    "\t\t{\"Weight\": %d, \"File\": \"data/motions/humanoid3d_%s.txt\"}" % (1 + (7 * k) % 5, "walk" if k % 3 else "run") for k in range(56)) + "\n\t]\n}\n"


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets.tar.gz")
    with tarfile.open(out, "w:gz", compresslevel=9) as tf:
        for f in FILES:
            p = os.path.join(REF, f)
            ti = tf.gettarinfo(p, arcname=f)
            ti.mtime = 0; ti.uid = ti.gid = 0; ti.uname = ti.gname = ""
            with open(p, "rb") as fh:
                tf.addfile(ti, fh)
        import io
        for name, text in ((MINI_DATASET, MINI_DATASET_TEXT), (SYN56_DATASET, SYN56_DATASET_TEXT)):
            data = text.encode()
            ti = tarfile.TarInfo(name)
            ti.size = len(data); ti.mtime = 0; ti.mode = 0o644
            tf.addfile(ti, io.BytesIO(data))
    print("wrote", out, os.path.getsize(out), "bytes")

if __name__ == "__main__":
    main()
