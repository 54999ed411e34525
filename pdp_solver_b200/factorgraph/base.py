"""Predict-side shell of the reference's FactorGraphTrainerBase (reference src/pdp/factorgraph/base.py:25-111,
254-305, 451-472).  Training, testing against labels and checkpoint management stay with the reference
(callers of the path, not the path); this class only drives `predict`."""
import os
import queue
import threading
import time

import torch

from .dataset import FactorGraphDataset


class FactorGraphTrainerBase(object):
    """Base class of the factor-graph predict pipeline.

    Multi-GPU: the DynamicBatchDivider segments of the input are independent work items -- every
    problem, and with `-b` every replica of it, lives in exactly one segment -- so the segments are handed out to the
    visible GPUs: one host process, one worker thread + one model replica + two CUDA streams per device (the next
    segment's host->device copy runs on the copy stream under the current segment's kernels), no data-path collective;
    the only gather is the per-problem output text, put back in input order.  Replaces the reference's nn.DataParallel
    wrapper (reference src/pdp/factorgraph/base.py:96-97), which would split graph_map[2,E] along dim 0.
    The random draws of a segment come from a generator seeded by (random_seed, segment index), so the output does not
    depend on how many GPUs shared the work."""

    def __init__(self, config, has_meta_data, error_dim, loss, evaluator, use_cuda, logger):
        self._config = config
        self._logger = logger
        if not use_cuda:
            raise RuntimeError("pdp_solver_b200 has no CPU path (--cpu_mode is the reference's own implementation)")
        if not torch.cuda.is_available():
            raise RuntimeError("pdp_solver_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._use_cuda = True
        first = torch.cuda.current_device()
        want = int(config.get("gpus") or os.environ.get("PDP_B200_GPUS", "0") or 0)
        count = torch.cuda.device_count()
        n = count if want <= 0 else min(want, count)
        self._devices = [torch.device("cuda", (first + i) % count) for i in range(max(n, 1))]
        self._device = self._devices[0]
        self._error_dim = error_dim
        self._loss = loss
        self._evaluator = evaluator
        if config.get("verbose"):
            self._logger.info("Using %d GPU(s): %s..." % (len(self._devices), torch.cuda.get_device_name(self._device)))
        # one model list per device; no nn.DataParallel
        self._models_by_device = []
        for dev in self._devices:
            self._device = dev
            with torch.cuda.device(dev):
                self._models_by_device.append([m.to(dev) for m in self._build_graph(self._config)])
        self._device = self._devices[0]
        self._model_list = self._models_by_device[0]

    def _build_graph(self, config):
        raise NotImplementedError("Subclass must implement abstract method")

    def _load(self, import_path_base):
        for models in self._models_by_device:
            for model in models:
                model.load(import_path_base)

    def _to_cuda(self, data):
        "base.py:100-106"
        if isinstance(data, list) or data is None:
            return data
        return data.cuda(self._device, non_blocking=True)

    def _check_recurrence_termination(self, active, prediction, sat_problem):
        pass

    # ---- base.py:254-305 ---------------------------------------------------------------------------
    def _segment_seed(self, k):
        return (int(self._config.get("random_seed", 0)) + 1000003 * k) % (2 ** 63)

    def _predict_epoch(self, batches, post_processor, batch_replication, file, segment_stream=None):
        def segments():
            if segment_stream is not None:        # lazily collated segments (FactorGraphDataset.segments)
                for k, seg in enumerate(segment_stream):
                    yield k, tuple(seg)
                return
            k = 0
            for data in batches:
                for i in range(len(data[0])):
                    yield k, tuple(d[i] for d in data)
                    k += 1

        out_lock = threading.Lock()
        done, state = {}, {"next": 0}

        def emit(k, message):       # output in input order, as soon as the next segment in line is there
            with out_lock:
                done[k] = message
                while state["next"] in done:
                    m = done.pop(state["next"])
                    if m is not None:
                        print(m, file=file)
                    state["next"] += 1

        if len(self._devices) == 1:
            self._device_worker(0, segments(), None, post_processor, batch_replication, emit)
            return
        feed = queue.Queue(maxsize=2 * len(self._devices))
        errors = []

        def run(slot):
            try:
                self._device_worker(slot, None, feed, post_processor, batch_replication, emit)
            except BaseException as e:    # surfaced by the main thread
                errors.append(e)

        threads = [threading.Thread(target=run, args=(i,), daemon=True) for i in range(len(self._devices))]
        for t in threads:
            t.start()
        try:
            for item in segments():
                while True:
                    if errors:
                        raise errors[0]
                    try:
                        feed.put(item, timeout=0.2)
                        break
                    except queue.Full:
                        pass
        finally:
            for _ in threads:
                while True:
                    try:
                        feed.put(None, timeout=0.2)
                        break
                    except queue.Full:
                        if errors:
                            break
            for t in threads:
                t.join()
        if errors:
            raise errors[0]

    def _device_worker(self, slot, it, feed, post_processor, batch_replication, emit):
        """Runs segments on device `slot`: the next segment's H2D copy is issued on the copy stream before the current
        segment's forward starts, so it overlaps the kernels."""
        dev = self._devices[slot]
        models = self._models_by_device[slot]
        torch.cuda.set_device(dev)
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)

        def fetch():
            if it is not None:
                return next(it, None)
            return feed.get()

        def upload(item):
            if item is None:
                return None
            k, host = item
            with torch.cuda.stream(copy):
                on_dev = [d if (isinstance(d, list) or d is None) else d.cuda(dev, non_blocking=True) for d in host]
                ev = torch.cuda.Event()
                ev.record(copy)
            return k, on_dev, ev

        with torch.no_grad():
            nxt = upload(fetch())
            while nxt is not None:
                k, tensors, ev = nxt
                nxt = upload(fetch())                 # in flight under this segment's kernels
                compute.wait_event(ev)
                for t in tensors:
                    if torch.is_tensor(t):
                        t.record_stream(compute)
                torch.cuda.default_generators[dev.index].manual_seed(self._segment_seed(k))
                message = self._predict_batch(models, *tensors, post_processor, batch_replication)
                emit(k, message)

    def _predict_batch(self, models, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat, label,
                       misc_data, post_processor, batch_replication):
        messages = []
        for model in models:
            state = model.get_init_state(graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                                         randomized=False, batch_replication=batch_replication)
            prediction, _ = model(
                init_state=state, graph_map=graph_map, batch_variable_map=batch_variable_map,
                batch_function_map=batch_function_map, edge_feature=edge_feature, meta_data=graph_feat,
                is_training=False, iteration_num=self._config["test_recurrence_num"],
                check_termination=self._check_recurrence_termination, batch_replication=batch_replication)
            if post_processor is not None and callable(post_processor):
                messages.append(post_processor(model, prediction, graph_map, batch_variable_map, batch_function_map,
                                               edge_feature, graph_feat, label, misc_data))
        return "\n".join(messages) if messages else None

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
        self._predict_epoch(None, post_processor, batch_replication, out_file,
                            segment_stream=dataset.segments(self._config["batch_size"], pin=True))
        for dev in self._devices:
            torch.cuda.synchronize(dev)
        duration = time.time() - start_time
        if self._config.get("verbose"):
            self._logger.info("Time spent: %s seconds" % duration)
        return duration
