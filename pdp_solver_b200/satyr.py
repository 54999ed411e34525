#!/usr/bin/env python3
"""Command line of the predict path: same positional arguments, options and output as the reference's
`src/satyr.py` (reference satyr.py:46-109), running the B200 solvers.

    python satyr.py <model_config.yaml> <test_path> <test_recurrence_num> [-b B] [-z BATCH] [-l LIMIT]
                    [-w LOCAL_SEARCH_ITERATIONS] [-e EPSILON] [-d] [-s SEED] [-o OUT.json] [-v]

Differences: `-c/--cpu_mode` is refused (the CPU implementation is the reference itself); with `-d` the DIMACS
files are scanned straight into batches (no temporary JSON file next to the input, satyr.py:74-86,106-107);
the YAML is read with `yaml.safe_load`.
"""
import argparse
import logging
import os
import sys
from datetime import datetime

import numpy as np
import torch
import yaml

from . import dimacs2json
from .trainer import SatFactorGraphTrainer


def build_parser():
    parser = argparse.ArgumentParser(prog="satyr.py", description="PDP SAT solver, predict path (B200)")
    parser.add_argument("model_config", help="The model configuration yaml file")
    parser.add_argument("test_path", help="The input test path")
    parser.add_argument("test_recurrence_num", help="The number of iterations for the PDP", type=int)
    parser.add_argument("-b", "--batch_replication", help="Batch replication factor", type=int, default=1)
    parser.add_argument("-z", "--batch_size", help="Batch size", type=int, default=5000)
    parser.add_argument("-m", "--max_cache_size", help="Maximum cache size", type=int, default=100000)
    parser.add_argument("-l", "--test_batch_limit", help="Memory limit for mini-batches", type=int, default=40000000)
    parser.add_argument("-w", "--local_search_iteration", help="Number of iterations for post-processing local search",
                        type=int, default=100)
    parser.add_argument("-e", "--epsilon", help="Epsilon probablity for post-processing local search", type=float,
                        default=0.5)
    parser.add_argument("-v", "--verbose", help="Verbose", action="store_true")
    parser.add_argument("-c", "--cpu_mode", help="Run on CPU (refused: use the reference for that)", action="store_true")
    parser.add_argument("-d", "--dimacs", help="The input folder contains DIMACS files", action="store_true")
    parser.add_argument("-s", "--random_seed", help="Random seed", type=int, default=int(datetime.now().microsecond))
    parser.add_argument("-o", "--output", help="The JSON output file", default="")
    parser.add_argument("-g", "--gpus", help="Number of GPUs the input segments are sharded over (default: all visible; "
                        "not a reference option)", type=int, default=0)
    return parser


def make_config(model_config, args):
    "satyr.py:88-101: merged configuration with the classical model types' overrides"
    config = {**model_config, **args}
    if config["model_type"] in ("p-d-p", "walk-sat", "reinforce"):
        config["model_path"] = None
        config["hidden_dim"] = 3
    if config["model_type"] == "walk-sat":
        config["local_search_iteration"] = config["test_recurrence_num"]
    config["dropout"] = 0
    config["error_dim"] = 1
    config["exploration"] = 0
    return config


def run(config, logger, output, rows=None):
    "Runs the prediction engine (satyr.py:20-43)."
    np.random.seed(config["random_seed"] % (2 ** 32))
    torch.manual_seed(config["random_seed"])
    if config["verbose"]:
        logger.info("Building the computational graph...")
    predicter = SatFactorGraphTrainer(config=config, use_cuda=not config["cpu_mode"], logger=logger)
    if config["verbose"]:
        logger.info("Starting the prediction phase...")
    predicter._counter = 0
    kw = dict(test_list=config["test_path"], import_path_base=config.get("model_path"),
              post_processor=predicter._post_process_predictions, batch_replication=config["batch_replication"], rows=rows)
    if output == "":
        predicter.predict(out_file=sys.stdout, **kw)
    else:
        with open(output, "w") as f:
            predicter.predict(out_file=f, **kw)
    return predicter


def main(argv=None):
    args = vars(build_parser().parse_args(argv))
    with open(args["model_config"], "r") as f:
        model_config = yaml.safe_load(f)
    logging.basicConfig(level=logging.DEBUG, format="[%(levelname)s] %(asctime)s - %(name)s: %(message)s")
    logger = logging.getLogger(model_config["model_name"])
    rows = None
    if args["dimacs"]:
        if args["verbose"]:
            logger.info("Scanning DIMACS files...")
        if os.path.isfile(args["test_path"]):
            rows = [dimacs2json.convert_one(args["test_path"], single_file=True)]
        else:
            rows = [dimacs2json.convert_one(p) for p in dimacs2json.dimacs_files(args["test_path"])]
    config = make_config(model_config, args)
    run(config, logger, config["output"], rows)
    print("")
    return 0


if __name__ == "__main__":
    sys.exit(main())
