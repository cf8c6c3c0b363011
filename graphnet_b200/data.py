"""Minimal `Data` / `Batch` containers with the `torch_geometric.data` contract.

The reference passes `torch_geometric.data.Data`/`Batch` objects between the
graph definition, the dataloader and `DynEdge.forward`
(`src/graphnet/models/graphs/graph_definition.py:205-247`,
`src/graphnet/data/dataloader.py:12-18`, `src/graphnet/models/gnn/dynedge.py:298`).
torch_geometric is not installed in this image, so the hot path is written
against the *fields* of that contract (`x`, `edge_index`, `batch`, `ptr`,
`n_pulses`, per-feature attributes) and accepts any object that has them -- a
real PyG `Batch` works unchanged.

The one extension is that `edge_index` may be backed by a device-resident
neighbour table produced by the B200 kNN kernel (`KnnGraph`); the int64
`[2, E]` tensor of the PyG contract is then materialised on first access only,
so the hot path never pays for it.
"""

from __future__ import annotations

from typing import Any, Dict, Iterable, Iterator, List, Optional

import torch


class Data:
    """Attribute bag with the subset of `torch_geometric.data.Data` used here."""

    def __init__(self, x: Optional[torch.Tensor] = None,
                 edge_index: Optional[torch.Tensor] = None, **kwargs: Any):
        object.__setattr__(self, "_store", {})
        object.__setattr__(self, "_knn", None)
        if x is not None:
            self._store["x"] = x
        if edge_index is not None:
            self._store["edge_index"] = edge_index
        for key, val in kwargs.items():
            self._store[key] = val

    # -- attribute / item protocol (PyG allows both) -------------------------
    def __getattr__(self, key: str) -> Any:
        if key.startswith("__"):
            raise AttributeError(key)
        store = object.__getattribute__(self, "_store")
        if key == "edge_index":
            return self._get_edge_index()
        if key in store:
            return store[key]
        if key in ("batch", "ptr"):
            return None
        raise AttributeError(key)

    def __setattr__(self, key: str, value: Any) -> None:
        if key == "edge_index":
            object.__setattr__(self, "_knn", None)
        self._store[key] = value

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def __setitem__(self, key: str, value: Any) -> None:
        setattr(self, key, value)

    def __contains__(self, key: str) -> bool:
        return key in self._store or (key == "edge_index" and self._knn is not None)

    def keys(self) -> List[str]:
        return list(self._store.keys())

    def _get_edge_index(self) -> Optional[torch.Tensor]:
        store = object.__getattribute__(self, "_store")
        if store.get("edge_index") is None and self._knn is not None:
            store["edge_index"] = self._knn.edge_index()
        return store.get("edge_index")

    # -- kernel-side graph cache ---------------------------------------------
    def set_knn_graph(self, graph: Any) -> None:
        """Attach a device neighbour table; `edge_index` becomes lazy."""
        self._store.pop("edge_index", None)
        object.__setattr__(self, "_knn", graph)

    def knn_graph(self) -> Any:
        return object.__getattribute__(self, "_knn")

    @property
    def num_nodes(self) -> int:
        return int(self._store["x"].shape[0])

    def to(self, device: Any, non_blocking: bool = False) -> "Data":
        out = self.__class__.__new__(self.__class__)
        object.__setattr__(out, "_store", {})
        object.__setattr__(out, "_knn", None)
        for key, val in self._store.items():
            out._store[key] = (val.to(device, non_blocking=non_blocking)
                               if isinstance(val, torch.Tensor) else val)
        if self._knn is not None:
            object.__setattr__(out, "_knn", self._knn.to(device))
        return out

    def __repr__(self) -> str:
        parts = []
        for key, val in self._store.items():
            if isinstance(val, torch.Tensor):
                parts.append(f"{key}={list(val.shape)}")
            else:
                parts.append(f"{key}={type(val).__name__}")
        return f"{self.__class__.__name__}({', '.join(parts)})"


class Batch(Data):
    """Several event graphs concatenated along the node dimension."""

    @classmethod
    def from_data_list(cls, data_list: Iterable[Data]) -> "Batch":
        """Collate like `torch_geometric.data.Batch.from_data_list`.

        Node-level tensors are concatenated, `edge_index` is offset by the
        cumulative node count, 0-dim tensors (`n_pulses`, scalar labels) are
        stacked to `[B]`, non-tensor attributes become lists, and `batch`/`ptr`
        are built (`src/graphnet/data/dataloader.py:12-18`).
        """
        data_list = list(data_list)
        assert len(data_list) > 0
        out = cls()
        sizes = [int(d.x.shape[0]) for d in data_list]
        offsets = [0]
        for s in sizes:
            offsets.append(offsets[-1] + s)
        keys = data_list[0].keys()
        for key in keys:
            vals = [d._store[key] for d in data_list]
            v0 = vals[0]
            if key == "edge_index":
                if any(v is None for v in vals):
                    continue
                out._store[key] = torch.cat(
                    [v + off for v, off in zip(vals, offsets[:-1])], dim=1)
            elif isinstance(v0, torch.Tensor):
                if v0.dim() == 0:
                    out._store[key] = torch.stack(vals)
                else:
                    out._store[key] = torch.cat(vals, dim=0)
            else:
                out._store[key] = vals
        if "edge_index" not in out._store and all(d.knn_graph() is not None for d in data_list):
            eis = [d.edge_index + off for d, off in zip(data_list, offsets[:-1])]
            out._store["edge_index"] = torch.cat(eis, dim=1)
        dev = data_list[0].x.device
        out._store["batch"] = torch.repeat_interleave(
            torch.arange(len(sizes), dtype=torch.int64, device=dev),
            torch.tensor(sizes, dtype=torch.int64, device=dev))
        out._store["ptr"] = torch.tensor(offsets, dtype=torch.int64, device=dev)
        return out

    @property
    def num_graphs(self) -> int:
        return int(self._store["ptr"].numel() - 1)

    def to_data_list(self) -> List[Data]:
        ptr = self._store["ptr"].tolist()
        ei = self.edge_index
        out = []
        for b in range(len(ptr) - 1):
            lo, hi = ptr[b], ptr[b + 1]
            d = Data()
            for key, val in self._store.items():
                if key in ("batch", "ptr", "edge_index"):
                    continue
                if isinstance(val, torch.Tensor):
                    if val.shape[0] == ptr[-1] and val.dim() >= 1 and key != "n_pulses":
                        d._store[key] = val[lo:hi]
                    else:
                        d._store[key] = val[b]
                elif isinstance(val, list):
                    d._store[key] = val[b]
            if ei is not None:
                m = (ei[1] >= lo) & (ei[1] < hi)
                d._store["edge_index"] = ei[:, m] - lo
            out.append(d)
        return out

    def __iter__(self) -> Iterator[Data]:
        return iter(self.to_data_list())
