"""Stub of configargparse on top of argparse: `is_config_file` arguments name a file of `key = value` lines whose
values become defaults (lists in [a, b] form), which is all the reference's config.py / configs/*.txt use."""
import argparse


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        k.pop('config_file_parser_class', None)
        k.pop('default_config_files', None)
        super().__init__(*a, **k)
        self._config_dests = []

    def add_argument(self, *a, **k):
        is_cfg = k.pop('is_config_file', False)
        act = super().add_argument(*a, **k)
        if is_cfg:
            self._config_dests.append(act.dest)
        return act

    def parse_known_args(self, args=None, namespace=None):
        ns, _ = super().parse_known_args(args, None)
        extra = []
        for dest in self._config_dests:
            path = getattr(ns, dest, None)
            if not path:
                continue
            with open(path) as f:
                for line in f:
                    line = line.split('#', 1)[0].strip()
                    if not line or '=' not in line:
                        continue
                    key, val = (s.strip() for s in line.split('=', 1))
                    if val.startswith('[') and val.endswith(']'):
                        vals = [v.strip() for v in val[1:-1].split(',') if v.strip()]
                    else:
                        vals = [val]
                    act = next((x for x in self._actions if x.dest == key or ('--' + key) in x.option_strings), None)
                    if act is None:
                        continue
                    if isinstance(act, (argparse._StoreTrueAction, argparse._StoreFalseAction)):
                        if vals[0].lower() in ('true', '1', 'yes'):
                            extra.append(act.option_strings[0])
                    else:
                        extra.append(act.option_strings[0])
                        extra.extend(vals)
        argv = extra + list(args if args is not None else __import__('sys').argv[1:])   # command line wins
        return super().parse_known_args(argv, namespace)
