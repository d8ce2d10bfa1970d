"""`key = value` configuration reader with the reference's semantics (R/AM_CommonTools/configuration/
configuration.py:3-45, 97-121): '#' starts a comment, keys are upper-cased, lines without exactly one '=' are
skipped, get() lets ast.literal_eval type the value.  Hot-path dependency only (not accelerated)."""
import ast


class Configuration:
    def __init__(self, config_data, key_order=None):
        self.data = config_data
        self.key_order = key_order

    @staticmethod
    def from_file(filename):
        data, order = {}, []
        with open(filename, "r") as f:
            for line in f:
                parts = line.split("#")[0].strip().split("=")
                if len(parts) != 2:
                    continue
                key = parts[0].strip().upper()
                data[key] = parts[1].strip()
                order.append(key)
        return Configuration(data, order)

    def get(self, name, default=None):
        if name not in self.data:
            return default
        try:
            return ast.literal_eval(self.data[name])
        except Exception:
            return self.data[name]

    def get_str(self, name, default=""):
        return self.data.get(name, default)

    def get_bool(self, name, default=False):
        return int(self.data[name]) > 0 if name in self.data else default

    def get_int(self, name, default=0):
        return int(self.data[name]) if name in self.data else default

    def get_float(self, name, default=0.0):
        return float(self.data[name]) if name in self.data else default

    def set(self, name, value):
        self.data[name] = value

    def contains(self, name):
        return name in self.data
