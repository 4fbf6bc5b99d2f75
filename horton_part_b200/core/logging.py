"""Logging helpers with the reference's behaviour (core/logging.py:27-102 of horton-part)."""

import logging
import os
import sys

__all__ = ["deflist", "setup_logger", "get_print_func"]


def deflist(logger, pairs):
    """Log ``name : value`` lines with the names padded to a common width."""
    width = max(len(name) for name, _ in pairs)
    for name, value in pairs:
        logger.info(f"  {name.ljust(width)} : {value}")


def setup_logger(logger, log_level=logging.INFO, log_file=None, overwrite=True):
    """(Re)configure ``logger`` with a single handler: a file if ``log_file`` else stdout."""
    if not isinstance(log_level, int):
        raise ValueError(f"Invalid log level: {log_level}")
    logger.setLevel(log_level)
    for old in list(logger.handlers):
        logger.removeHandler(old)
    fmt = "%(levelname)s: %(message)s" if log_level <= logging.DEBUG else "%(message)s"
    if log_file:
        folder = os.path.dirname(log_file)
        if folder:
            os.makedirs(folder, exist_ok=True)
        append = os.path.exists(log_file) and not overwrite
        handler = logging.FileHandler(log_file, mode="a" if append else "w")
    else:
        handler = logging.StreamHandler(sys.stdout)
    handler.setFormatter(logging.Formatter(fmt))
    logger.addHandler(handler)


def get_print_func(logger=None, verbose=False):
    if logger is None:
        return print
    return logger.info if verbose else logger.debug
