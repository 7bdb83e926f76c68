"""Token-level similarity (difflib ratio after stripping comments and docstrings) between a file of this package and the
reference file of the same role.  Run in the build container only (needs /root/reference)."""
import difflib
import io
import sys
import tokenize


def tokens(path):
    out = []
    prev_type = tokenize.INDENT
    with open(path, "rb") as fh:
        for tok in tokenize.tokenize(fh.readline):
            if tok.type in (tokenize.COMMENT, tokenize.NL, tokenize.NEWLINE, tokenize.ENCODING, tokenize.INDENT, tokenize.DEDENT, tokenize.ENDMARKER):
                prev_type = tok.type if tok.type in (tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT) else prev_type
                continue
            if tok.type == tokenize.STRING and prev_type in (tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT):
                continue      # docstring / bare string statement
            prev_type = tok.type
            out.append(tok.string)
    return out


if __name__ == "__main__":
    a, b = tokens(sys.argv[1]), tokens(sys.argv[2])
    sm = difflib.SequenceMatcher(None, a, b, autojunk=False)
    matched = sum(bl.size for bl in sm.get_matching_blocks())
    print(f"ratio {sm.ratio():.3f}  matched {matched}  of ours {len(a)} ({matched / len(a):.2f})  of ref {len(b)} ({matched / len(b):.2f})")
