"""nnet.config reader -- same contract as /root/reference/nnet/config.py:40-63: one `key ... value`
pair per line (first whitespace token is the key, LAST token the value), `#` starts a comment, and
values are coerced int -> float -> bool ('true'/'false', any case) -> str, in that order."""


def _coerce(text):
    for conv in (int, float):
        try:
            return conv(text)
        except ValueError:
            pass
    low = text.lower()
    if low == "true":
        return True
    if low == "false":
        return False
    return text


def parse_config(fn):
    config = {}
    with open(fn, "r") as fh:
        for raw in fh:
            line = raw.strip()
            if line.startswith("#"):
                continue
            fields = [tok for tok in line.split() if not tok.startswith("#")]
            # the reference indexes tokens[0] unconditionally (IndexError on a blank line); we skip blanks
            if not fields:
                continue
            config[fields[0]] = _coerce(fields[-1])
    return config
