"""`import caffe2.python._import_c_extension as C` — the reference only reads
C.num_cuda_devices (models/dlrm_s_caffe2.py:103) and calls C.benchmark_net."""
num_cuda_devices = 0


def benchmark_net(name, warmup_runs, main_runs, run_individual):
    from caffe2.python import workspace
    for _ in range(int(warmup_runs) + int(main_runs)):
        workspace.RunNet(name)
    return []
