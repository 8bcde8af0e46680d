"""TF-style global flags over argparse.

Drop-in for the reference's config layer (/root/reference/core/flags.py:14-135):
`FLAGS.<name>` parses lazily on first access, unknown arguments are ignored,
booleans accept `--flag`, `--flag=True|False`, `--flag False` (argparse
nargs='?': execute.py passes `--enable_network_costs False`) and `--noflag`.
"""
from __future__ import annotations

import argparse

_parser = argparse.ArgumentParser(description="gsched-b200 simulator flags")


class _Flags:
    def __init__(self):
        object.__setattr__(self, "_values", {})
        object.__setattr__(self, "_parsed", False)

    def _parse(self, args=None):
        ns, rest = _parser.parse_known_args(args=args)
        self._values.update(vars(ns))
        object.__setattr__(self, "_parsed", True)
        return rest

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if not self._parsed:
            self._parse()
        try:
            return self._values[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        if not self._parsed:
            self._parse()
        self._values[name] = value

    def reset(self, args=None):
        """Re-parse (tests and in-process sweeps)."""
        self._values.clear()
        return self._parse(args)


FLAGS = _Flags()


def _define(name, default, doc, kind):
    _parser.add_argument("--" + name, default=default, help=doc, type=kind)


def DEFINE_string(name, default, doc):
    _define(name, default, doc, str)


def DEFINE_integer(name, default, doc):
    _define(name, default, doc, int)


def DEFINE_float(name, default, doc):
    _define(name, default, doc, float)


def DEFINE_boolean(name, default, doc):
    def as_bool(text):
        return str(text).lower() in ("true", "t", "1")
    _parser.add_argument("--" + name, nargs="?", const=True, default=default,
                         type=as_bool, help=doc)
    _parser.add_argument("--no" + name, action="store_false", dest=name)


DEFINE_bool = DEFINE_boolean


def DEFINE_version(text):
    _parser.add_argument("-v", "--version", action="version", version="%(prog)s " + text)


_defined = False


def define_simulator_flags():
    """The flag surface of /root/reference/run_sim.py:19-94 (names, types, defaults)."""
    global _defined
    if _defined:
        return FLAGS
    _defined = True
    import time
    DEFINE_string("trace_file", "tf_job.csv", "job trace CSV (live schema)")
    DEFINE_string("log_path", "result-" + time.strftime("%Y%m%d-%H-%M-%S", time.localtime()),
                  "output folder under ./log/")
    DEFINE_string("scheme", "yarn", "placement scheme: yarn | count | horus | horus+ | gandiva")
    DEFINE_string("schedule", "fifo", "policy: fifo | sjf | dlas | dlas-gpu | gittins | horus | horus+ | gandiva")
    DEFINE_boolean("pack", False, "pack several tasks per GPU (not supported by the engine)")
    DEFINE_integer("num_switch", 1, "switches in the cluster")
    DEFINE_integer("num_node_p_switch", 32, "nodes under one switch")
    DEFINE_boolean("enable_network_costs", False, "add PS<->worker transfer time after placement")
    DEFINE_boolean("enable_migration", False, "accepted for CLI compatibility; no effect on fifo")
    DEFINE_integer("bandwidth", 1250, "rack bandwidth, MB/s")
    DEFINE_float("internode_latency", 0.015, "latency per crossed node, seconds")
    DEFINE_integer("gpu_memory_capacity", 32, "GPU memory, GiB")
    DEFINE_integer("num_queue", 1, "queues in the job manager (dlas MLFQ depth; horus+ credit queues)")
    DEFINE_integer("num_buffer", 5, "look-ahead width of the horus scheduler")
    DEFINE_integer("num_gpu_p_node", 8, "GPUs per node")
    DEFINE_integer("num_cpu_p_node", 128, "CPUs per node")
    DEFINE_integer("mem_p_node", 512, "memory per node")
    DEFINE_string("cluster_spec", None, "CSV overriding the five topology flags")
    DEFINE_boolean("print", False, "accepted for CLI compatibility")
    DEFINE_boolean("flush_stdout", True, "accepted for CLI compatibility")
    # engine-side additions (absent from the reference; defaults keep its behaviour)
    DEFINE_integer("device", 0, "CUDA device ordinal")
    DEFINE_string("trace_cache", "log/.trace_cache", "directory for parsed traces (sweeps replay one trace many times); empty: always parse")
    DEFINE_string("queue_limit", "3600,7200,18000", "dlas thresholds, comma separated")
    DEFINE_float("gittins_delta", 3250.0, "gittins service quantum")
    DEFINE_version("0.1")
    return FLAGS
