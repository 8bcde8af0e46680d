"""Job-side host objects: thin mirrors of JobQueueManager / JobsManager / Job
(/root/reference/core/jobs/job_queue_manager.py:9-29, jobs_manager.py:9-26,
job.py:60-110) that hold the admission-ordered JobTable the device consumes."""
from __future__ import annotations

from . import ingest

SUPPORTED_SCHEDULES = ("fifo", "sjf", "dlas", "dlas-gpu", "gittins", "horus", "horus+", "gandiva")


class JobQueueManager:
    def __init__(self, flags, file_path=None):
        self.flags = flags
        self.file_path = file_path
        self.num_queue = flags.num_queue


class Job:
    """What placement / network-cost entry points need to know about one job."""

    def __init__(self, job_id, duration, submit_time, gpu_p_worker=1, gpu_memory_max=0, total_gpus=1,
                 ps_count=0, model_size=0.0, iterations=0.0):
        self.job_id = str(job_id)
        self.duration = duration
        self.submit_time = int(submit_time)
        self.gpu_per_worker = gpu_p_worker
        self.gpus = total_gpus
        self.task_count = int(total_gpus // gpu_p_worker)
        self.gpu_mem_max = gpu_memory_max            # MiB, like the reference
        self.ps_count = ps_count
        self.model_size = model_size
        self.iterations = iterations
        self.tasks_running_on = {}
        self.held = []                               # [(node index, device mask, tasks)] set by the placement call
        self.start_time = 0
        self.end_time = 0
        self.migration_count = 0
        self._processed = 0

    def is_distributed(self):
        return self.ps_count > 1

    def add_network_costs(self, extra_s):
        self.duration += extra_s

    def get_duration(self):
        return self.duration

    def time_processed(self):
        return self._processed


class JobsManager:
    def __init__(self, flags, job_queue_manager):
        self.job_queue_manager = job_queue_manager
        self.flags = flags
        if flags.schedule not in SUPPORTED_SCHEDULES:
            # the reference dies at the first insert (jobs_manager.py:62)
            raise NotImplementedError(flags.schedule)
        self.replay_trace = True
        self._reader = None
        self.table = ingest.load_table(flags.trace_file, 0.5,        # scale factor: schedule.py:187
                                       getattr(flags, "trace_cache", None) or None)
        self.running_jobs = {}
        self.finished_jobs = {}

        self.queue = []                              # queue 0 of the reference, head = index 0
        self._next_row = 0

    @property
    def job_generator(self):
        """the reference's reader object (jobs_manager.py:16-18); built on first use -- a run served from the parsed-trace
        cache never needs the frame"""
        if self._reader is None:
            self._reader = ingest.JobTraceReader(self.flags.trace_file).prepare_jobs()
        return self._reader

    def remaining_jobs(self, delta_time=None):
        return self.table.n - self._next_row

    # ---- per-call surface of the reference's JobsManager (jobs_manager.py:32-37,115-140,228-241), host mirror:
    # the fused engine never calls these; they exist so that the registry entries of algorithm.py can be driven tick by
    # tick like the reference drives them (tests/test_gpu_parity.py::test_per_call_registries_step_like_the_fused_loop)
    def gen_jobs(self, delta_time, scale_factor=0.5):
        """admit every trace row whose arrival tick is due; the batch lands at the HEAD of the queue (quirk Q2)"""
        t = self.table
        batch = []
        while self._next_row < t.n and t.arrive_tick[self._next_row] <= delta_time:
            j = self._next_row
            job = Job(t.label[j] if t.label else j, float(t.duration[j]), int(t.submit[j]), int(t.gpu_per_task[j]),
                      float(t.mem_bytes[j]) / 1048576.0, int(t.gpus[j]))
            job.index = j
            batch.append(job)
            self._next_row += 1
        self.insert(batch)
        return batch

    def insert(self, jobs):
        jobs = jobs if isinstance(jobs, list) else [jobs]
        self.queue[0:0] = jobs                       # positions 0..n-1 (jobs_manager.py:52-55,134-135)

    def get_next_job(self, delta_time):
        return self.queue[0] if self.queue and self.queue[0].submit_time <= delta_time else None

    def pop(self, delta_time):
        return self.queue.pop(0)

    def queuing_jobs(self, delta_time=None):
        return len(self.queue)
