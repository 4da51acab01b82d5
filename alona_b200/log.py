"""Logging helpers with the reference's call shape (alona/log.py:15-54):
log_error() logs and exits with status 1, exactly like the reference."""
import logging
import sys


def init_logging(loglevel='regular', logfile='alona.log'):
    level = logging.DEBUG if loglevel == 'debug' else logging.INFO
    logging.basicConfig(filename=logfile, level=level, format='%(asctime)s %(message)s')
    console = logging.StreamHandler()
    console.setLevel(level)
    console.setFormatter(logging.Formatter('%(levelname)-8s %(message)s'))
    logging.getLogger('').addHandler(console)


def log_info(msg):
    logging.info('%s', msg)


def log_debug(msg):
    logging.debug('%s', msg)


def log_warning(msg):
    logging.warning('%s', msg)


def log_error(msg='error'):
    logging.error('%s', msg)
    sys.exit(1)
