"""CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT.

fp32 PyTorch restatements of the reference's scoring hot path (zuokai/KDDCUP_2020_MultimodalitiesRecall_2nd_Place),
each function citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package, and only as the checker or the timed CPU baseline.
The product (kddcup_2020_multimodalitiesrecall_2nd_place_b200) never imports it and has no CPU fallback.

Pinning status
  * ensemble (code/main.py) and nDCG@5: pinned by the reference's shipped files — tests/golden/ensemble_kat and
    the nDCG known answer 0.7098 (kdd-report-final.pdf table 5) reproduced from shipped score files — and nDCG@k
    also by the reference's own evaluation.py run on seeded synthetic rankings (tests/golden/ndcg_ref_cases.json).
  * LXMERT: pinned against the reference's OWN PyTorch code, imported from /root/reference in the dev container
    (oracle/lxmert_ref.py); golden vectors generated from it are committed under tests/golden/ with the script.
  * ImageBert zk / lds: TF-1.12 + Python 2 are not installable, and the reference ships no activations or weights.
    Pinned at the level of the model code: the reference's OWN graph-building code (imagebert_zk/pixelbert.py +
    model_triple.py, imagebert_lds/src/pixelmodel.py + get_next_sentence_output) runs unmodified on tools/tf1_shim.py,
    an eager stand-in for the TensorFlow ops it calls, and its outputs on seeded synthetic weights / inputs are
    committed as tests/golden/{zk,lds}_ref_shim_*.npz (tools/make_golden.py --tf-shim; re-derived from
    /root/reference on every CPU test run in the dev container).  PARITY UNPINNED only below that: the arithmetic of
    the individual TensorFlow ops is restated in the shim from the TF 1.12 documentation.
"""
