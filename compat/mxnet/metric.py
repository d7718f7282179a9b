"""`eval_metric='accuracy'` of Module.fit (train.py:82)."""


class Accuracy:
    name = "accuracy"

    def __init__(self):
        self.reset()

    def reset(self):
        self.num_inst, self.sum_metric = 0, 0.0

    def update(self, labels, preds):
        for label, pred in zip(labels, preds):
            p = pred.asnumpy().argmax(axis=1)
            y = label.asnumpy().astype("int64")
            self.sum_metric += float((p == y).sum())
            self.num_inst += len(y)

    def get(self):
        return self.name, (self.sum_metric / self.num_inst if self.num_inst else float("nan"))


def create(metric):
    if isinstance(metric, str) and metric in ("acc", "accuracy"):
        return Accuracy()
    if hasattr(metric, "update"):
        return metric
    raise NotImplementedError("metric %r (the reference uses 'accuracy')" % (metric,))
