from zbot_lab_b200.utils.configclass import configclass  # noqa: F401
