"""world_size-2 gloo run of the outcome sharding used by GPSearch.penalized_optimization (no GPU needed: only the
partition and the host-side result exchange are exercised; the fit itself is replaced by a stub)."""
import os
import socket
import sys

import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from waveome_b200 import datasets, model_search
    from waveome_b200.model_search import GPSearch

    def fake_fit(X, Y, models, **kw):          # stands in for the CUDA engine
        B = len(models)
        for m in models:
            m.fit_info = dict(status=0)
        return dict(x=np.zeros((B, 1)), f=np.arange(B, dtype=float), lml=np.zeros(B), n_iter=np.zeros(B, np.int32),
                    n_eval=np.ones(B, np.int32), status=np.zeros(B, np.int32))
    model_search.fit_models = fake_fit

    def fake_replicated(X, Y, template, make_models=None, **kw):
        models = make_models()
        return fake_fit(X, Y, models), models
    model_search.fit_replicated = fake_replicated
    model_search.feature_importances_batch = lambda X, Y, models, **kw: [[0.0, 1.0] for _ in models]
    X, Y = datasets.overview_synthetic(n_people=6, n_observations=4, n_outcomes=7)
    g = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
    g.penalized_optimization()
    n_pen = g.fit_report["n_models"]
    pen_names = sorted(g.models.keys())
    # run_search: the lock-step search of this rank's shard with a stub fitter, then the same exchange
    from waveome_b200 import kernel_search as ks

    def fake_search_fit(requests):
        return [(ks.candidate_model(k), float(len(name))) for _y, name, k in requests]
    g.run_search(max_depth=2, fit=fake_search_fit)
    assert sorted(g.models.keys()) == pen_names and all(v["best_model"] for v in g.search_info.values())
    q.put((rank, pen_names, n_pen))
    dist.destroy_process_group()


def test_outcomes_are_sharded_and_gathered():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    names = [f"outcome{j + 1}" for j in range(7)]
    assert res[0][1] == sorted(names) and res[1][1] == sorted(names)      # every rank ends with every model
    assert res[0][2] + res[1][2] == 7 and abs(res[0][2] - res[1][2]) <= 1   # disjoint, balanced shards
