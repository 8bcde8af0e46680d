"""Cluster description + live resource tables as flat arrays.

Mirror of the reference's cluster model (/root/reference/infra/infrastructure.py:16-147,
infra/node.py:7-34): same constructor (`Infrastructure(flags)`), same attributes the
hot path reads (`nodes`, `num_gpu_p_node`, `bandwidth`, `internode_latency`,
`enable_network_costs`, `gpu_memory_capacity` in MiB) and the same spec-file
override rules, but the per-node state is a structure of arrays
(busy-device mask, cpu_used, mem_used) -- the layout the device kernels consume.
"""
from __future__ import annotations

import csv
import os
from collections import OrderedDict

import numpy as np

from . import capi

SPEC_KEYS = ["num_switch", "num_node_p_switch", "num_gpu_p_node", "num_cpu_p_node", "mem_p_node"]


class NodeView:
    """Read/write view of one row of the node table (Node.node_id is 1-based, as a string)."""

    def __init__(self, infra, index):
        self._i = infra
        self.index = index
        self.node_id = str(index + 1)
        self.rack_id = str(index // infra.num_nodes_p_switch)

    @property
    def cpu_used(self):
        return int(self._i.table["cpu_used"][self.index])

    @property
    def mem_used(self):
        return int(self._i.table["mem_used"][self.index])

    def cpu_free(self):
        return self._i.num_cpu_p_node - self.cpu_used

    def mem_free(self):
        return self._i.mem_p_node - self.mem_used

    def is_free(self):                       # node.py:59-60
        return self.cpu_free() > 0 or self.mem_free() > 0

    def get_free_devices(self, pack=False):  # node.py:99-107
        mask = int(self._i.table["busy_mask"][self.index])
        return self._i.num_gpu_p_node - bin(mask).count("1")


class Infrastructure:
    def __init__(self, flags):
        self.flags = flags
        self.num_switch = flags.num_switch
        self.bandwidth = flags.bandwidth
        self.internode_latency = flags.internode_latency
        self.enable_network_costs = flags.enable_network_costs
        self.gpu_memory_capacity = flags.gpu_memory_capacity * 1024      # MiB (infrastructure.py:36)
        self.num_nodes_p_switch = flags.num_node_p_switch
        self.num_cpu_p_node = flags.num_cpu_p_node
        self.num_gpu_p_node = flags.num_gpu_p_node
        self.mem_p_node = flags.mem_p_node
        self.cluster_spec = getattr(flags, "cluster_spec", None)
        # --pack: the reference stores it on every Node / Device (infrastructure.py:55, node.py:22, device.py:12) and
        # never reads it again -- packing is decided by the placement scheme (pack=True inside horus_placement) -- so
        # the flag is accepted and, as there, changes nothing.
        self.enable_pack = bool(getattr(flags, "pack", False))
        if self.cluster_spec and os.path.exists(self.cluster_spec):
            self._init_from_spec_file()
        self._init_nodes()

    def _init_from_spec_file(self):
        # same lookup rule as infrastructure.py:78-81: relative to the package root's parent
        project_dir = os.path.abspath(os.path.dirname(os.path.dirname(__file__)))
        spec_file = os.path.join(project_dir, self.cluster_spec)
        _, ext = os.path.splitext(spec_file)
        assert "csv" in ext
        with open(spec_file, "r") as fh:
            reader = csv.DictReader(fh, delimiter=",")
            keys = reader.fieldnames
            for k in SPEC_KEYS:
                if k not in keys:
                    return
            for row in reader:
                self.num_switch = int(row["num_switch"])
                self.num_nodes_p_switch = int(row["num_node_p_switch"])
                self.num_gpu_p_node = int(row["num_gpu_p_node"])
                self.num_cpu_p_node = int(row["num_cpu_p_node"])
                self.mem_p_node = int(row["mem_p_node"])

    def _init_nodes(self):
        m = self.num_switch * self.num_nodes_p_switch
        self.table = np.zeros(m, dtype=capi.NODE_DTYPE)
        self.nodes = OrderedDict((str(i + 1), NodeView(self, i)) for i in range(m))
        self.racks = OrderedDict((str(r), [self.nodes[str(r * self.num_nodes_p_switch + k + 1)]
                                           for k in range(self.num_nodes_p_switch)])
                                 for r in range(self.num_switch))

    def get_total_gpus(self):
        return len(self.nodes) * self.num_gpu_p_node

    def get_free_nodes(self):
        return [n for n in self.nodes.values() if n.is_free()]

    def gs_cluster(self) -> capi.GsCluster:
        return capi.make_cluster(self.num_switch, self.num_nodes_p_switch, self.num_gpu_p_node,
                                 self.num_cpu_p_node, self.mem_p_node, self.flags.gpu_memory_capacity,
                                 self.enable_network_costs, self.bandwidth, self.internode_latency)
