"""Generates tests/golden/reference_glue.json by RUNNING the reference's own glue functions.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_reference_glue_golden.py
The reference's modules cannot be imported (they import dwave.* at module level), so the function bodies are extracted
from its source with ``ast`` and executed here; only their OUTPUTS are committed:
  * greedy_get_subgraph / get_graph_mapping (src/utils/common.py:22-100): selected node lists, sub-graph edges and
    physical -> logical mappings on Zephyr Z4 and Pegasus P4 for several seeds and sizes;
  * get_latent_to_discrete("heaviside") (src/utils/common.py:143-175): spins and straight-through gradient of a fixed
    logits tensor;
  * train_grbm (src/model_wrapper.py:59-67): the update schedule's truth table;
  * the first five logged losses of the shipped training runs (models/*/losses.json; data, not code).
tests/test_reference_glue_golden.py compares this repository's restatements with these outputs on any machine.
"""
import ast
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/src"


def extract(path, names, namespace):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), path, "exec"), namespace)
    return namespace


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found; fixtures are generated in the build container only")
    import networkx as nx
    import torch

    import image_generation_b200 as B

    ns = extract(os.path.join(REF, "utils", "common.py"), {"greedy_get_subgraph", "get_graph_mapping", "get_latent_to_discrete"},
                 {"random": random, "nx": nx, "DWaveSampler": None, "torch": torch})
    out = {"subgraphs": []}
    for topo, size in (("zephyr", 4), ("pegasus", 4)):
        n, ei, ej, _ = (B.zephyr_graph if topo == "zephyr" else B.pegasus_graph)(size)
        graph = nx.Graph()
        graph.add_nodes_from(range(n))
        graph.add_edges_from(zip(ei.tolist(), ej.tolist()))
        for seed, k in ((775321899904, 64), (1, 17), (12345, 128), (7, 256)):
            if k > n:
                continue
            sub = ns["greedy_get_subgraph"](n_nodes=k, random_seed=seed, graph=graph)
            _, mapping = ns["get_graph_mapping"](sub)
            out["subgraphs"].append({"topology": topo, "size": size, "seed": seed, "n_nodes": k,
                                     "nodes": [int(v) for v in sub.nodes()],
                                     "edges": sorted([min(int(a), int(b)), max(int(a), int(b))] for a, b in sub.edges()),
                                     "mapping": [[int(a), int(b)] for a, b in mapping.items()]})
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(7, 33, generator=g)
    x = logits.clone().requires_grad_(True)
    spins = ns["get_latent_to_discrete"]("heaviside")(x, 1)
    (spins * torch.arange(33.0)).sum().backward()
    out["heaviside"] = {"logits": logits.tolist(), "spins": spins.detach().tolist(), "grad": x.grad.tolist(),
                        "none_mode_is_none": ns["get_latent_to_discrete"](None) is None}
    mw = extract(os.path.join(REF, "model_wrapper.py"), {"train_grbm"}, {})
    out["train_grbm"] = [[bool(mw["train_grbm"](step, epoch)) for epoch in range(9)] for step in range(40)]
    # first logged losses of the six shipped training runs (models/*/losses.json): dvae_loss - mse_loss at step 0 is the
    # MMD term of a randomly initialised model -- the only reference-computed MMD values in the tree
    out["first_step_losses"] = {}
    models = os.path.join(os.path.dirname(REF), "models")
    for name in sorted(os.listdir(models)):
        d = json.load(open(os.path.join(models, name, "losses.json")))
        out["first_step_losses"][name] = {"mse": d["mse_losses"][:5], "dvae": d["dvae_losses"][:5]}
    path = os.path.join(HERE, "reference_glue.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
