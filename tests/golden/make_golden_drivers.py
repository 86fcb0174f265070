"""Generates tests/golden/drivers_small.npz: the reference's own EVALUATION DRIVER `node_infer_Gs_GD` (run.py:49-115)
called unmodified — behind oracle/ref_shims.py — on the subgraph lists, split masks and checkpoints already stored in
node_small.npz (node classification: NLLLoss_numpy + accuracy) and node_reg_small.npz (node regression:
L1Loss_numpy / std(labels)), for the 'test' and 'val' branches, both loss reductions and all three modes; and the TRAINING
DRIVER `node_train_Gs_GD` (run.py:177-215): three Adam steps on node_small.npz, dropout disabled (see below).

    python tests/golden/make_golden_drivers.py          # authoring container only
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference's utils / network)
from oracle import ref_shims  # noqa: E402
from tests import golden_io as gio  # noqa: E402

_cwd = os.getcwd()
import tempfile  # noqa: E402
os.chdir(tempfile.mkdtemp())  # run.py creates ./results on import
import run as ref_run  # noqa: E402  (reference run.py)
os.chdir(_cwd)


def graphs_of(d, mode):
    out = []
    for s in gio.subgraphs(d, mode + "_sub"):
        g = mg.Data(x=torch.tensor(s["x"]), edge_index=torch.tensor(s["edge_index"]), y=torch.tensor(s["y"]))
        g.train_mask, g.val_mask, g.test_mask = (torch.tensor(s[k]) for k in ("train_mask", "val_mask", "test_mask"))
        out.append(g)
    return out


def main():
    out = {}
    for case, task, Model, Loss in (("node_small", "node_cls", mg.ref_network.Classify_node, mg.ref_utils.NLLLoss_numpy),
                                    ("node_reg_small", "node_reg", mg.ref_network.Regress_node, mg.ref_utils.L1Loss_numpy)):
        d = gio.load(case)
        sd = gio.state_dict(d)
        C = int(d["n_classes"]) if "n_classes" in d.files else 1
        for mode in ("none", "extra", "cluster"):
            graphs = graphs_of(d, mode)
            for reduction in ("mean", "sum"):
                args = argparse.Namespace(task=task, num_classes=C, num_features=d["x"].shape[1], hidden=int(d["hidden"]),
                                          num_layers1=2, layer_name="GCNConv", loss_reduction=reduction)
                model = Model(args)
                model.load_state_dict(sd)
                loader = ref_shims.DataLoader(graphs, batch_size=128, shuffle=False)  # run.py:336
                for which in ("test", "val"):
                    with torch.no_grad():
                        loss, acc, _ = ref_run.node_infer_Gs_GD(args, model, loader, Loss(reduction), which)
                    out[f"{case}_{mode}_{reduction}_{which}"] = np.array([loss, acc], dtype=np.float64)
    # ---- mini-batch evaluation driver: node_infer_Gs_MB (run.py:117-175): per-batch torch losses, averaged over ALL batches
    for case, task, Model, Loss in (("node_small", "node_cls", mg.ref_network.Classify_node, torch.nn.NLLLoss),
                                    ("node_reg_small", "node_reg", mg.ref_network.Regress_node, torch.nn.L1Loss)):
        d = gio.load(case)
        C = int(d["n_classes"]) if "n_classes" in d.files else 1
        for mode in ("none", "extra", "cluster"):
            graphs = graphs_of(d, mode)
            for reduction in ("mean", "sum"):
                args = argparse.Namespace(task=task, num_classes=C, num_features=d["x"].shape[1], hidden=int(d["hidden"]),
                                          num_layers1=2, layer_name="GCNConv", loss_reduction=reduction)
                model = Model(args)
                model.load_state_dict(gio.state_dict(d))
                loader = ref_shims.DataLoader(graphs, batch_size=32, shuffle=False)  # several batches on the small lists
                for which in ("test", "val"):
                    with torch.no_grad():
                        loss, acc, _ = ref_run.node_infer_Gs_MB(args, model, loader, Loss(reduction=reduction), which)
                    out[f"mb_{case}_{mode}_{reduction}_{which}"] = np.array([loss, acc], dtype=np.float64)
    out["mb_batch_size"] = np.array(32)

    # ---- graph-level evaluation driver: graph_infer_Gs (run.py:306-328) over T_DataLoader(..., collate_fn=colater())
    # batches (run.py:577-580, :709-712), with the loss functions graph_classification / graph_regression construct
    # (CrossEntropyLoss on the model's softmax output, run.py:583; L1Loss, run.py:716)
    from torch.utils.data import DataLoader as T_DataLoader
    for case, task, Model, loss_fn in (("graph_small", "graph_reg", mg.ref_network.Regress_graph_gs, torch.nn.L1Loss()),
                                       ("graph_cls_small", "graph_cls", mg.ref_network.Classify_graph_gs,
                                        torch.nn.CrossEntropyLoss())):
        d = gio.load(case)
        data_list = []
        for g in range(int(d["n_kept"])):
            subs = [mg.Data(x=torch.tensor(s_["x"]), edge_index=torch.tensor(s_["edge_index"]), mask=torch.tensor(s_["mask"]))
                    for s_ in gio.subgraphs(d, f"g{g}_sub")]
            graph = mg.Data(x=torch.tensor(d[f"g{g}_x"]), edge_index=torch.tensor(d[f"g{g}_ei"]), y=torch.tensor(d[f"g{g}_y"]))
            Gc = mg.Data(x=torch.tensor(d[f"g{g}_gc_x"]), edge_index=torch.tensor(d[f"g{g}_gc_edge"]))
            data_list.append([graph, Gc, subs])
        args = argparse.Namespace(task=task, num_classes=int(d["n_classes"]) if "n_classes" in d.files else 1,
                                  num_features=d["g0_x"].shape[1], hidden=int(d["hidden"]), num_layers1=2,
                                  layer_name="GCNConv", multi_prop=False)
        model = Model(args)
        model.load_state_dict(gio.state_dict(d))
        loader = T_DataLoader(data_list, batch_size=3, shuffle=False, collate_fn=mg.ref_utils.colater())
        import warnings
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")  # L1Loss broadcasts [B,1] against [B] in the reference (see infer.graph_metrics)
            loss, acc = ref_run.graph_infer_Gs(args, model, loader, loss_fn)
        out[f"graph_{case}"] = np.array([loss, acc], dtype=np.float64)
    out["graph_batch_size"] = np.array(3)

    # ---- the inference script's own model classes Net1 / Net2 (inference.py:72-116).  inference.py is a script (argparse and
    # dataset downloads at module level), so the two class definitions are cut out with `ast` and executed with the names
    # they use (GCNConv = the shim the whole fixture set uses, F, torch); the checkpoints of Classify_node / Regress_node load
    # into them by key (inference.py:668-670), and the per-sample loop (inference.py:672-688) reads row j of one subgraph.
    import ast
    import torch.nn.functional as F_inf
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "inference.py")).read()
    ns = {"torch": torch, "F": F_inf, "GCNConv": sys.modules["torch_geometric.nn"].GCNConv}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name in ("Net1", "Net2"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "inference.py", "exec"), ns)
    d = gio.load("node_small")
    net1 = ns["Net1"](d["x"].shape[1], int(d["hidden"]), 2, int(d["n_classes"]))
    net1.load_state_dict(gio.state_dict(d))  # strict: identical keys
    net1.eval()
    subs = gio.subgraphs(d, "extra_sub")
    with torch.no_grad():
        out["net1_per_query"] = np.stack([net1(torch.tensor(subs[i]["x"]), torch.tensor(subs[i]["edge_index"]))[j].numpy()
                                          for i, j in ((0, 0), (3, 1), (17, 0), (40, 2), (98, 0))])
    out["net1_queries"] = np.array([(0, 0), (3, 1), (17, 0), (40, 2), (98, 0)])
    d = gio.load("node_reg_small")
    net2 = ns["Net2"](d["x"].shape[1], int(d["hidden"]), 2)
    net2.load_state_dict(gio.state_dict(d))
    net2.eval()
    subs = gio.subgraphs(d, "cluster_sub")
    with torch.no_grad():
        out["net2_per_query"] = np.stack([net2(torch.tensor(subs[i]["x"]), torch.tensor(subs[i]["edge_index"]))[j].numpy()
                                          for i, j in ((0, 0), (5, 1), (20, 0), (60, 1))])
    out["net2_queries"] = np.array([(0, 0), (5, 1), (20, 0), (60, 1)])

    # ---- training driver: node_train_Gs_GD (run.py:177-215), three Adam steps (lr / weight decay: main.py defaults).
    # The dropout mask of network.py:33 depends on the RNG stream, so the step is recorded with F.dropout replaced by the
    # identity for the duration of the call (torch.nn.functional is patched, no reference source is touched).
    d = gio.load("node_small")
    import torch.nn.functional as F_
    real_dropout = F_.dropout
    F_.dropout = lambda x, p=0.5, training=True, inplace=False: x
    try:
        for mode in ("none", "extra", "cluster"):
            args = argparse.Namespace(task="node_cls", num_classes=int(d["n_classes"]), num_features=d["x"].shape[1],
                                      hidden=int(d["hidden"]), num_layers1=2, layer_name="GCNConv", loss_reduction="mean")
            model = mg.ref_network.Classify_node(args)
            model.load_state_dict(gio.state_dict(d))
            opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
            loader = ref_shims.DataLoader(graphs_of(d, mode), batch_size=128, shuffle=False)
            losses = [ref_run.node_train_Gs_GD(model, loader, torch.nn.NLLLoss(reduction="mean"), opt, args) for _ in range(3)]
            out[f"train_{mode}_losses"] = np.array(losses, dtype=np.float64)
            for k, v in model.state_dict().items():
                out[f"train_{mode}_sd_{k}"] = v.detach().numpy()
        # node_train_Gc / node_val_Gc (run.py:26-47): the same model trained on the COARSENED graph Gc of the fixture
        # (load_data_classification outputs stored in node_small.npz), three Adam steps + the validation loss after them
        for mode in ("none", "cluster"):
            args = argparse.Namespace(task="node_cls", num_classes=int(d["n_classes"]), num_features=d["x"].shape[1],
                                      hidden=int(d["hidden"]), num_layers1=2, layer_name="GCNConv", loss_reduction="mean")
            model = mg.ref_network.Classify_node(args)
            model.load_state_dict(gio.state_dict(d))
            opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
            gx, ge = torch.tensor(d[f"{mode}_gc_x"]), torch.tensor(d[f"{mode}_gc_edge"])
            ty, tm = torch.tensor(d[f"{mode}_gc_train_y"]), torch.tensor(d[f"{mode}_gc_train_m"])
            vy, vm = torch.tensor(d[f"{mode}_gc_val_y"]), torch.tensor(d[f"{mode}_gc_val_m"])
            lf = torch.nn.NLLLoss(reduction="mean")
            losses = [ref_run.node_train_Gc(model, gx, ge, tm, ty, lf, opt) for _ in range(3)]
            with torch.no_grad():
                vloss = ref_run.node_val_Gc(model, gx, ge, vm, vy, lf)
            out[f"train_gc_{mode}_losses"] = np.array(losses + [vloss], dtype=np.float64)
            for k, v in model.state_dict().items():
                out[f"train_gc_{mode}_sd_{k}"] = v.detach().numpy()
    finally:
        F_.dropout = real_dropout
    np.savez_compressed(os.path.join(mg.OUT, "drivers_small.npz"), **out)
    for k in sorted(out)[:6]:
        print(k, out[k])
    print(len(out), "entries")


if __name__ == "__main__":
    main()
