"""Minimal stand-in for harmonic/logs.py: the hot path only needs warning_log
(harmonic/evidence.py:382,392)."""
import logging

logger = logging.getLogger("Harmonic")


def debug_log(message):
    logger.debug(message)


def info_log(message):
    logger.info(message)


def warning_log(message):
    logger.warning(message)


def critical_log(message):
    logger.critical(message)
