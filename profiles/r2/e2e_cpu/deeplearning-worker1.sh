source /etc/profile

CUDA_VISIBLE_DEVICES='' python /root/repo/deeplearning_cfn_b200/examples/cifar10_multi_machine_train.py --train_dir /tmp/dlcfn-e2e-cpu/stacks/e2e/nodes/i-d0e87274/myEFSvolume/cifar10_train --ps_hosts=deeplearning-worker1:2222,deeplearning-worker2:2222,deeplearning-worker3:2222,deeplearning-worker4:2222 --worker_hosts=deeplearning-worker1:2230,deeplearning-worker2:2230,deeplearning-worker3:2230,deeplearning-worker4:2230 --job_name=ps --task_index=0 > /tmp/dlcfn-e2e-cpu/stacks/e2e/nodes/i-d0e87274/myEFSvolume/logs/ps0 2>&1 &

CUDA_VISIBLE_DEVICES='0' python /root/repo/deeplearning_cfn_b200/examples/cifar10_multi_machine_train.py --train_dir /tmp/dlcfn-e2e-cpu/stacks/e2e/nodes/i-d0e87274/myEFSvolume/cifar10_train --ps_hosts=deeplearning-worker1:2222,deeplearning-worker2:2222,deeplearning-worker3:2222,deeplearning-worker4:2222 --worker_hosts=deeplearning-worker1:2230,deeplearning-worker2:2230,deeplearning-worker3:2230,deeplearning-worker4:2230 --job_name=worker --task_index=0 > /tmp/dlcfn-e2e-cpu/stacks/e2e/nodes/i-d0e87274/myEFSvolume/logs/worker0 2>&1 &

