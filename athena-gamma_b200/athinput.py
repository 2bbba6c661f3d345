"""ParameterInput: the reference's run-time configuration surface
(src/parameter_input.{hpp,cpp}: athinput files of `<block>` / `key = value # comment` lines,
`block/key=value` command-line overrides that may only modify EXISTING keys,
Get{Integer,Real,Boolean,String} / GetOrAdd*)."""
import re


class ParameterInput:
    def __init__(self, text=None, path=None):
        self.blocks = {}
        if path is not None:
            with open(path) as f:
                text = f.read()
        if text:
            self.load(text)

    def load(self, text):
        cur = None
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            m = re.match(r"<(\w+)>", line)
            if m:
                cur = m.group(1)
                if cur == "par_end":
                    break
                self.blocks.setdefault(cur, {})
                continue
            if cur is None:
                raise ValueError("### FATAL ERROR in ParameterInput: parameter outside a block: " + raw)
            if "=" not in line:
                raise ValueError("### FATAL ERROR in ParameterInput: no '=' in line: " + raw)
            k, v = line.split("=", 1)
            self.blocks[cur][k.strip()] = v.strip()

    def modify_from_cmdline(self, args):
        """src/parameter_input.cpp:351: block and key must already exist."""
        for a in args:
            m = re.match(r"(\w+)/(\w+)=(.*)", a)
            if not m:
                raise ValueError("### FATAL ERROR in ModifyFromCmdline: bad override " + a)
            b, k, v = m.groups()
            if b not in self.blocks:
                raise KeyError("### FATAL ERROR in ModifyFromCmdline: Block name '%s' not found" % b)
            if k not in self.blocks[b]:
                raise KeyError("### FATAL ERROR in ModifyFromCmdline: Parameter '%s/%s' not found" % (b, k))
            self.blocks[b][k] = v

    def does_parameter_exist(self, block, key):
        return block in self.blocks and key in self.blocks[block]

    def _get(self, block, key):
        if not self.does_parameter_exist(block, key):
            raise KeyError("### FATAL ERROR in ParameterInput: Parameter name '%s' not found in "
                           "block '%s'" % (key, block))
        return self.blocks[block][key]

    def get_integer(self, block, key):
        return int(self._get(block, key))

    def get_real(self, block, key):
        return float(self._get(block, key))

    def get_string(self, block, key):
        return self._get(block, key)

    def get_boolean(self, block, key):
        v = self._get(block, key).lower()
        return v in ("1", "true", "t", "yes")

    def _get_or_add(self, block, key, default, conv):
        if self.does_parameter_exist(block, key):
            return conv(self.blocks[block][key])
        self.blocks.setdefault(block, {})[key] = str(default)
        return default

    def get_or_add_integer(self, block, key, default):
        return self._get_or_add(block, key, default, int)

    def get_or_add_real(self, block, key, default):
        return self._get_or_add(block, key, default, float)

    def get_or_add_string(self, block, key, default):
        return self._get_or_add(block, key, default, str)

    def get_or_add_boolean(self, block, key, default):
        return self._get_or_add(block, key, default, lambda v: v.lower() in ("1", "true", "t", "yes"))

    def set(self, block, key, value):
        self.blocks.setdefault(block, {})[key] = str(value)

    def dump(self):
        out = []
        for b, kv in self.blocks.items():
            out.append("<%s>" % b)
            out += ["%s = %s" % (k, v) for k, v in kv.items()]
            out.append("")
        return "\n".join(out)
