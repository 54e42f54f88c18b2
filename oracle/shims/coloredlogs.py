"""No-op stand-in for the third-party `coloredlogs` package (absent offline).
Test infrastructure only: lets /root/reference import so goldens can be generated
(reference use: utils/logger.py:1,4,26)."""
import logging


class ColoredFormatter(logging.Formatter):
    pass


def install(**kwargs):
    return None
