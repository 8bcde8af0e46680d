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
        self.job_generator = ingest.JobTraceReader(flags.trace_file)
        self.job_generator.prepare_jobs()
        self.replay_trace = True
        self.table = self.job_generator.table(scale_factor=0.5)      # schedule.py:187
        self.running_jobs = {}
        self.finished_jobs = {}

    def remaining_jobs(self, delta_time=None):
        return self.table.n
