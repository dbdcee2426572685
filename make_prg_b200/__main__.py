"""`python -m make_prg_b200 from_msa ...` -- same common flags as the reference CLI
(make_prg/__main__.py:13-96); `-t` is accepted for compatibility (the locus pool is replaced by GPU
batching)."""
import argparse
import os
import sys

from loguru import logger

from .subcommands import from_msa
from .subcommands.output_type import OutputType


def setup_logger(verbose, log_file):
    logger.remove()
    level = {0: "INFO", 1: "DEBUG"}.get(verbose, "TRACE")
    logger.add(log_file if log_file else sys.stderr, level=level, enqueue=True,
               format="{time:YYYY-MM-DD HH:mm:ss} | {level} | {message}")


def main(argv=None):
    parser = argparse.ArgumentParser(prog="make_prg",
                                     description="Subcommand entrypoint (B200-native from_msa core)")
    subparsers = parser.add_subparsers(title="Available subcommands", dest="command")
    subparsers.required = True
    p = from_msa.register_parser(subparsers)
    p.add_argument("-O", "--output-type", dest="output_type", action="store", default="a", type=OutputType,
                   help="p: PRG, b: Binary, g: GFA, a: All. Combinations are allowed. Default: a")
    p.add_argument("-F", "--force", action="store_true", dest="force", help="Force overwrite previous output")
    p.add_argument("-t", "--threads", action="store", type=int, default=1,
                   help="Accepted for compatibility with the reference CLI")
    p.add_argument("-v", "--verbose", action="count", default=0)
    p.add_argument("--log", dest="log", action="store", type=str, default=None)
    options = parser.parse_args(argv)
    setup_logger(options.verbose, options.log)
    if options.threads == 0:
        options.threads = os.cpu_count()
    options.func(options)


if __name__ == "__main__":
    main()
