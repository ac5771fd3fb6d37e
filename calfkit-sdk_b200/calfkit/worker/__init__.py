from calfkit.worker.worker import Worker

__all__ = ["Worker"]
