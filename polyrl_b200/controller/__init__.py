from .decorator import Dispatch, Execute, register, make_nd_compute_dataproto_dispatch_fn
from .worker_group import (ClassWithInitArgs, LocalWorkerGroup, MPWorkerGroup, RayWorkerGroup, Worker, WorkerGroup,
                           create_colocated_worker_cls, have_ray)
