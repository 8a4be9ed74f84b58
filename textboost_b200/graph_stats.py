"""Exact launch accounting of a captured step: the kernel nodes of its CUDA graph, by kernel name.

``bench.py`` reports ``gpu_launches`` (how many of this library's kernels run inside the timed region).  The ctypes
binding can only estimate that (an entry point such as tb_groupnorm_fwd_f16 launches one or two kernels depending on the
path the C side picks), so the step is captured once more into a graph that is kept (``torch.cuda.CUDAGraph(keep_graph=
True)``), never instantiated or replayed, and its nodes are read back through cuda-python: every kernel node, with the
function name the driver reports.  Names in the ``tb`` namespace are this library's; the rest are torch's (a few fills and
one add) and, under data parallelism, NCCL's all-reduce.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch


def kernel_nodes(enqueue: Callable[[], object]) -> Dict[str, object]:
    """Capture ``enqueue()`` on a side stream into a kept CUDA graph and count its kernel nodes.  Nothing executes.
    Returns {"kernel_nodes", "tb_kernels", "other_kernels": {name: count}, "memset_nodes", "by_name": {name: count}}."""
    from cuda.bindings import driver as drv

    g = torch.cuda.CUDAGraph(keep_graph=True)
    with torch.cuda.graph(g):
        enqueue()
    graph = drv.CUgraph(int(g.raw_cuda_graph()))
    err, _, n = drv.cuGraphGetNodes(graph, 0)
    if err != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"cuGraphGetNodes: {err}")
    err, nodes, n = drv.cuGraphGetNodes(graph, n)
    if err != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"cuGraphGetNodes: {err}")
    by_name: Dict[str, int] = {}
    kernels = memsets = 0
    for node in nodes[:n]:
        err, kind = drv.cuGraphNodeGetType(node)
        if err != drv.CUresult.CUDA_SUCCESS:
            continue
        if kind == drv.CUgraphNodeType.CU_GRAPH_NODE_TYPE_MEMSET:
            memsets += 1
        if kind != drv.CUgraphNodeType.CU_GRAPH_NODE_TYPE_KERNEL:
            continue
        kernels += 1
        name = "?"
        err, params = drv.cuGraphKernelNodeGetParams(node)
        if err == drv.CUresult.CUDA_SUCCESS:
            err, raw = drv.cuFuncGetName(params.func)
            if err == drv.CUresult.CUDA_SUCCESS and raw:
                name = raw.decode() if isinstance(raw, (bytes, bytearray)) else str(raw)
        by_name[name] = by_name.get(name, 0) + 1
    ours = {k: v for k, v in by_name.items() if _is_ours(k)}
    other = {k: v for k, v in by_name.items() if not _is_ours(k)}
    del g
    return {"kernel_nodes": kernels, "tb_kernels": sum(ours.values()), "other_kernels": other, "memset_nodes": memsets,
            "by_name": by_name}


def _is_ours(name: str) -> bool:
    # mangled names of namespace tb start with _ZN2tb (functions) -- templates included; demangled ones with tb::
    return name.startswith("_ZN2tb") or "tb::" in name
