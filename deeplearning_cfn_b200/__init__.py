"""deeplearning_cfn_b200 - a B200-native distributed-training launcher + runtime with the capabilities of
awslabs/deeplearning-cfn (cluster provisioning / rendezvous / hostfile + env contract, launchers, KVStore
parameter server, Horovod-style all-reduce, example workloads), rebuilt for one NVSwitch-connected box of
Blackwell GPUs: PyTorch for the substrate, hand-written sm_100a kernels for the communication-bound hot
paths, NCCL kept as the baseline."""
__version__ = "0.1.0"
