"""Predict-side shell of the reference's FactorGraphTrainerBase (reference src/pdp/factorgraph/base.py:25-111,
254-305, 451-472).  Training, testing against labels and checkpoint management stay with the reference
(callers of the path, not the path); this class only drives `predict`."""
import time

import torch

from .dataset import FactorGraphDataset


class FactorGraphTrainerBase(object):
    "Base class of the factor-graph predict pipeline."

    def __init__(self, config, has_meta_data, error_dim, loss, evaluator, use_cuda, logger):
        self._config = config
        self._logger = logger
        if not use_cuda:
            raise RuntimeError("pdp_solver_b200 has no CPU path (--cpu_mode is the reference's own implementation)")
        if not torch.cuda.is_available():
            raise RuntimeError("pdp_solver_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._use_cuda = True
        self._device = torch.device("cuda", torch.cuda.current_device())
        self._error_dim = error_dim
        self._loss = loss
        self._evaluator = evaluator
        if config.get("verbose"):
            self._logger.info("Using GPU %s..." % torch.cuda.get_device_name(self._device))
        # one model per entry; no nn.DataParallel (it would split graph_map[2,E] along dim 0)
        self._model_list = [m.to(self._device) for m in self._build_graph(self._config)]

    def _build_graph(self, config):
        raise NotImplementedError("Subclass must implement abstract method")

    def _load(self, import_path_base):
        for model in self._model_list:
            model.load(import_path_base)

    def _to_cuda(self, data):
        "base.py:100-106"
        if isinstance(data, list) or data is None:
            return data
        return data.cuda(self._device, non_blocking=True)

    def _check_recurrence_termination(self, active, prediction, sat_problem):
        pass

    # ---- base.py:254-305 ---------------------------------------------------------------------------
    def _predict_epoch(self, batches, post_processor, batch_replication, file):
        with torch.no_grad():
            for data in batches:
                for i in range(len(data[0])):
                    (graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat, label,
                     misc_data) = [self._to_cuda(d[i]) for d in data]
                    self._predict_batch(graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                                        label, misc_data, post_processor, batch_replication, file)

    def _predict_batch(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat, label,
                       misc_data, post_processor, batch_replication, file):
        for model in self._model_list:
            state = model.get_init_state(graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                                         randomized=False, batch_replication=batch_replication)
            prediction, _ = model(
                init_state=state, graph_map=graph_map, batch_variable_map=batch_variable_map,
                batch_function_map=batch_function_map, edge_feature=edge_feature, meta_data=graph_feat,
                is_training=False, iteration_num=self._config["test_recurrence_num"],
                check_termination=self._check_recurrence_termination, batch_replication=batch_replication)
            if post_processor is not None and callable(post_processor):
                message = post_processor(model, prediction, graph_map, batch_variable_map, batch_function_map,
                                         edge_feature, graph_feat, label, misc_data)
                print(message, file=file)

    # ---- base.py:451-472 ---------------------------------------------------------------------------
    def predict(self, test_list, out_file, import_path_base=None, post_processor=None, batch_replication=1, rows=None):
        """Produces predictions.  `rows`: already-parsed problems (the DIMACS input of satyr.py -d) instead of a
        JSON file."""
        start_time = time.time()      # the input is scanned when the dataset is built: part of the time reported
        dataset = FactorGraphDataset(
            input_file=test_list, limit=self._config["test_batch_limit"], hidden_dim=self._config["hidden_dim"],
            max_cache_size=self._config.get("max_cache_size", 100000), batch_replication=batch_replication, rows=rows)
        if import_path_base is not None:
            self._load(import_path_base)
        self._predict_epoch(dataset.batches(self._config["batch_size"], pin=True), post_processor, batch_replication,
                            out_file)
        torch.cuda.synchronize(self._device)
        duration = time.time() - start_time
        if self._config.get("verbose"):
            self._logger.info("Time spent: %s seconds" % duration)
        return duration
