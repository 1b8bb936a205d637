"""Reference-named entry points of the scoring hot path (same module paths, function names, argument order and
shapes as /code in the reference), backed by the sm_100a kernels through the C ABI.

TF-1 graph construction + `sess.run` (imagebert_zk, imagebert_lds) becomes an eager call on numpy / torch arrays;
`saver.restore` / `load_state_dict` becomes `bind(weights)` with the checkpoint's own variable names.  Inference only:
`is_training=True` raises.  Nothing here computes on the CPU: every function ends in libmmrecall.so kernels and
raises when no sm_100 device is present.
"""
