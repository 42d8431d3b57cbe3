"""shim of ``omni`` (train.py:80,150): only ``omni.log.warn`` is used."""
import warnings


class log:  # noqa: N801
    @staticmethod
    def warn(msg):
        warnings.warn(str(msg))

    @staticmethod
    def info(msg):
        pass
