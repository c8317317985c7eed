"""Debug: per-row difference between the persistent kernel and the per-kernel path (mid-sized config)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_persistent_gpu as t
for batch in [int(a) for a in sys.argv[1:]] or [16, 17, 33]:
    cfg = t._mid_config(batch)
    P, R = cfg.max_prefill_predict_length, cfg.max_target_length - cfg.max_prefill_predict_length
    pl, al = t._ragged(batch, P, R, seed=batch)
    outs = []
    for persistent in (True, False):
        engine, dparams = t._engine(cfg, persistent)
        state = engine.fill_synthetic_context(pl, al, seed=11)
        state, _ = engine.generate(dparams, state)
        outs.append(state["logits"].float().cpu().clone())
    d = (outs[0] - outs[1]).abs().amax(dim=(1, 2))
    print("batch", batch, "per-row max diff:", np.round(d.numpy(), 3).tolist())
