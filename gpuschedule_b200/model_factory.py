"""Checkpoint sizes (MB) keyed by model name.

Constant table consumed by the PS<->worker transfer-time model: it is the
source of `job.model_size` in
/root/reference/core/network/network_service.py:34.  The values restate the
table at /root/reference/model/model_factory.py:19-55 (a published list of
checkpoint sizes); nothing else of that module is on the hot path.
"""

model_sizes = dict(
    [("4_layers_brnn", 1300), ("transformer", 1100),
     ("1_layer_bilstm_opennmt", 900), ("BERT_Chinese", 350),
     ("2_layers_lstm_gigaword", 330), ("mobilenet_v1_025", 15),
     ("googlenet", 26), ("inception2", 43), ("inception3", 104),
     ("inception4", 176), ("alexnet", 233), ("vgg11", 519), ("vgg19", 549),
     ("vgg16", 528), ("resnet50", 97), ("resnet101", 555),
     ("resnet152", 737)])


def model_size_mb(name, default=0.0):
    """MB for `name`; unknown names cost nothing (no transfer term)."""
    return float(model_sizes.get(str(name), default))
