"""The train step replayed from a CUDA graph (Engine(graph=True)) takes the same optimisation steps as the kernel-by-kernel
launch sequence: the Adam step size comes from the device-side step counter, so replay k uses beta powers of step k."""
import numpy as np
import pytest
import torch

from oracle import pamrec_oracle as O

pytestmark = pytest.mark.gpu


def test_graph_replay_matches_eager_steps():
    from pamrec_b200.engine import Engine
    nu, ni, nc, T, B = 300, 3000, 50, 50, 100
    om = O.OracleModel(nu, ni, nc, T, seed=5)
    O.perturb_params(om.params, om.bn_state, seed=6)
    engs = [Engine(nu, ni, nc, T, B, graph=g).allocate() for g in (False, True)]
    for e in engs:
        e.set_variables({n: t.numpy() for n, t in om.params.items()})
        e.set_variables({n: t.numpy() for n, t in om.bn_state.items()})
    batches = [O.make_batch(40 + i, B, T, nu, ni, nc) for i in range(2)]
    dbs = [[e.upload(b) for b in batches] for e in engs]
    got = [[], []]
    for step in range(7):
        for k, e in enumerate(engs):
            got[k].append(e.train_step(dbs[k][step % 2]).cpu().numpy().copy())
    eager, graphed = engs
    assert len(graphed._graphs) == 2 and not eager._graphs            # first use of a batch eager, then one graph per resident batch
    assert graphed.step == eager.step == 7
    assert float(graphed.ws("adam.step")[0]) == 7.0
    for step in range(7):
        assert np.allclose(got[0][step], got[1][step], rtol=2e-5, atol=1e-7), (step, got[0][step], got[1][step])
    a, b = eager.get_variables(), graphed.get_variables()
    lr = eager.hp["learning_rate"]
    for name in a:
        # the two runs differ by the order of fp32 atomics only; Adam turns that noise into at most ~lr per step on near-zero gradients
        assert np.abs(a[name] - b[name]).max() <= 2.5 * lr, name
    # the losses also track the fp64 oracle like the eager steps do
    for step in range(7):
        ref = om.train_step(batches[step % 2])["losses"]
        assert abs(got[1][step][0] - ref["loss"]) <= 1e-4 * max(abs(ref["loss"]), 1e-3), (step, got[1][step][0], ref["loss"])
    # a resumed optimiser step must not be replayed blindly: the next step runs eagerly and re-seeds the device counter
    graphed.set_optimizer_state({"step": 20})
    graphed.train_step(dbs[1][0])
    torch.cuda.synchronize()
    assert graphed.step == 21 and float(graphed.ws("adam.step")[0]) == 21.0
    graphed.train_step(dbs[1][1])
    torch.cuda.synchronize()
    assert float(graphed.ws("adam.step")[0]) == 22.0
    for e in engs:
        e.close()
