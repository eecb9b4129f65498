"""CPU: host-side helpers of the evaluation drivers and of the bench's CPU arm."""
import os
import subprocess
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_images_do_not_depend_on_the_process(tmp_path):
    """ImageSource's stand-in noise images must be the same bytes in every process (they were seeded with hash(name), which Python salts
    per process: the GPU test that compares batched and one-by-one decoding then saw different inputs on every run)."""
    code = ("import sys; sys.path.insert(0, %r); from evaluation.common import ImageSource; from PIL import Image; import numpy as np, zlib; "
            "p = ImageSource('/nonexistent', %r).path('COCO_val2014_000000000042.jpg', (97, 61)); "
            "print(zlib.crc32(np.asarray(Image.open(p)).tobytes()))")
    outs = []
    for i, seed in enumerate(("1", "2")):
        d = tmp_path / f"s{i}"
        env = dict(os.environ, PYTHONHASHSEED=seed)
        outs.append(subprocess.run([sys.executable, "-c", code % (REPO, str(d))], capture_output=True, text=True, env=env, check=True).stdout.strip())
    assert outs[0] == outs[1] and outs[0].isdigit()
    from evaluation.common import ImageSource
    from PIL import Image
    p = ImageSource("/nonexistent", str(tmp_path / "s2")).path("COCO_val2014_000000000042.jpg", (97, 61))
    img = np.asarray(Image.open(p))
    assert img.shape == (61, 97, 3) and str(zlib.crc32(img.tobytes())) == outs[0]


def test_usable_cores_respects_affinity_and_cap():
    sys.path.insert(0, REPO)
    from oracle.reference_path import usable_cores
    n = usable_cores()
    assert 1 <= n <= 32 and n <= (os.cpu_count() or 1)
    assert usable_cores(cap=1) == 1
    try:
        assert n <= len(os.sched_getaffinity(0))
    except AttributeError:
        pass
