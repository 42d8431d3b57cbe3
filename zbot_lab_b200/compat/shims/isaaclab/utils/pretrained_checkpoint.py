def get_published_pretrained_checkpoint(workflow: str, task_name: str):
    """play.py:64, 106: there are no published checkpoints for the ZBOT tasks."""
    return None
