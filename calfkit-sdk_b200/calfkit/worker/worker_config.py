from dataclasses import dataclass


@dataclass
class WorkerConfig:
    """declared by the reference (calfkit/worker/worker_config.py:7-14) and unused there as well"""
    max_workers: int = 1
    group_id: str | None = None
