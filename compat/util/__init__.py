"""`from util import Visulizer` (reference util/visulization.py:4-48).  The reference's constructor dials a visdom server
(http://hpc3.yud.io:8088); this stand-in keeps the interface and writes to the logging module instead."""
import logging
import time


class Visulizer(object):
    def __init__(self, host="http://hpc3.yud.io", port=8088, env="street"):
        self.host, self.port, self.env = host, port, env
        self.index = {}
        self.log_text = ""

    def reinit(self, env="default"):
        return self

    def plot(self, name, y):
        x = self.index.get(name, 0)
        logging.info("[vis %s] %s[%d] = %s", self.env, name, x, y)
        self.index[name] = x + 1

    def img(self, name, img_, **kwargs):
        pass

    def log(self, info, win="log_text"):
        self.log_text += "[{time}] {info} <br>".format(time=time.strftime("%m-%d %H:%M:%S"), info=info)
        logging.info("[vis %s] %s", self.env, info)

    def delete_env(self, env):
        pass
