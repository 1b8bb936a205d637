"""Record-decode throughput: the multi-threaded C++ decoder (mmr_decode_tsv) against the reference's per-line Python /
numpy decode (imagebert_zk/load_data_v4.py:133-147) on synthetic TSV lines of the competition's shape."""
import base64
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import records  # noqa: E402

rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lines = []
for i in range(n):
    nb = int(np.clip(rng.poisson(4) + 1, 1, 10))
    boxes = (rng.random((nb, 4)) * 500).astype(np.float32)
    feats = rng.standard_normal((nb, 2048)).astype(np.float32)
    labels = rng.integers(0, 33, nb).astype(np.int64)
    f = [str(i), "600", "800", str(nb), base64.b64encode(boxes.tobytes()).decode(), base64.b64encode(feats.tobytes()).decode(),
         base64.b64encode(labels.tobytes()).decode(), "women leather shoes", str(i // 30)]
    lines.append(("\t".join(f) + "\n").encode())
mb = sum(map(len, lines)) / 1e6


def reference_decode(line):
    arr = line.decode().strip().split("\t")
    nb = int(arr[3])
    b = np.frombuffer(base64.b64decode(arr[4]), dtype=np.float32).reshape(nb, 4)
    b5 = np.zeros((nb, 5), dtype=np.float32)
    b5[:, :4] = b / [int(arr[1]), int(arr[2]), int(arr[1]), int(arr[2])]
    f = np.frombuffer(base64.b64decode(arr[5]), dtype=np.float32).reshape(nb, 2048)
    c = np.frombuffer(base64.b64decode(arr[6]), dtype=np.int64).reshape(nb)
    return b5, np.concatenate([f, np.zeros((10 - nb, 2048))]), c


t0 = time.perf_counter()
for ln in lines:
    reference_decode(ln)
t_ref = time.perf_counter() - t0
print(f"reference-style python decode : {n / t_ref:9.0f} lines/s  {mb / t_ref:8.1f} MB/s (1 thread)")
for th in (1, 4, 0):
    dec = records.RecordDecoder(n, n_threads=th, pin=False)
    dec.decode(lines)                      # first touch of the buffers
    t0 = time.perf_counter()
    for _ in range(3):
        dec.decode(lines)
    dt = (time.perf_counter() - t0) / 3
    print(f"mmr_decode_tsv threads={th or os.cpu_count():<3d}      : {n / dt:9.0f} lines/s  {mb / dt:8.1f} MB/s")
