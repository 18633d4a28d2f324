"""Asset root resolution.  DeepMimic arg files name assets relative to the directory that contains
data/ and args/ (the reference's repository root).  Order: $DEEPMIMIC_ASSET_ROOT, /root/reference when it
exists (build container), else the archive tests/golden/assets.tar.gz unpacked once next to it."""
import os
import tarfile
import warnings
import threading

_lock = threading.Lock()
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def asset_root(prefer_archive: bool = False) -> str:
    env = os.environ.get("DEEPMIMIC_ASSET_ROOT")
    if env:
        return env
    if not prefer_archive and os.path.isdir("/root/reference/data/characters"):
        return "/root/reference"
    arc = os.path.join(_REPO, "tests", "golden", "assets.tar.gz")
    out = os.path.join(_REPO, "tests", "golden", "_assets")
    with _lock:
        stamp = os.path.join(out, ".unpacked")
        if not os.path.exists(stamp) or os.path.getmtime(stamp) < os.path.getmtime(arc):
            os.makedirs(out, exist_ok=True)
            with tarfile.open(arc, "r:gz") as tf:
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    tf.extractall(out)
            open(stamp, "w").close()
    return out
