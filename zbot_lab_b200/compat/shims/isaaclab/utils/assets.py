import os


def retrieve_file_path(path: str, download_dir=None, force_download: bool = True) -> str:
    """``isaaclab.utils.assets.retrieve_file_path`` (play.py:62, 111): local files only (no Nucleus, no network)."""
    if os.path.isfile(path):
        return os.path.abspath(path)
    raise FileNotFoundError(f"Unable to find the file: {path}")
