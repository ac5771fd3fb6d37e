from calfkit.engine.batch import BatchEngine, BatchOutput, Publish, ToolTemplate

__all__ = ["BatchEngine", "BatchOutput", "Publish", "ToolTemplate"]
