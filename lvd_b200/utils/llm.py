"""Layout-response parsing: the part of the reference's utils/llm.py the generation loop calls with a cached response
(`get_parsed_layout(prompt, override_response=resp, json_template=...)`, generate.py:305-310; utils/llm.py:169-328).
Querying an LLM is out of scope (no network): without `override_response` these raise.
"""
import ast
import json

REQUIRED_LINES = [f"Frame {i + 1}:" for i in range(6)] + ["Background keyword:"]  # prompt.py:65-66
REQUIRED_LINES_AST = [True] * 6 + [False]
_STRIP = " \t\n`"


def get_parsed_layout(*args, json_template=False, **kwargs):
    return (get_parsed_layout_json_resp if json_template else get_parsed_layout_text_resp)(*args, **kwargs)


def get_parsed_layout_text_resp(prompt, llm_kwargs=None, max_partial_response_retries=1, override_response=None,
                                strip_chars=_STRIP, save_leading_text=True, **kwargs):
    """Text template: `Frame 1: [...]` ... `Frame 6: [...]` `Background keyword: ...`.  Returns
    (parsed_layout, reconstructed_response) like utils/llm.py:176-274."""
    if override_response is None:
        raise NotImplementedError("lvd_b200 does not query an LLM: pass override_response (a cached response)")
    assert max_partial_response_retries == 1, "override_response is specified so no partial queries are allowed"
    text = override_response
    rebuilt = ""
    head = REQUIRED_LINES[0]
    if head in text:
        parts = text.split(head)
        if save_leading_text:
            rebuilt += parts[0] + head
        text = parts[1]
    layout = {}
    prev_cut = 0
    for i, line in enumerate(REQUIRED_LINES):
        nxt = REQUIRED_LINES[i + 1] if i + 1 < len(REQUIRED_LINES) else ""
        if nxt not in text:
            raise ValueError(f"Erroring due to many non-successful attempts on prompt: {prompt} with response "
                             f"{override_response}")
        cut = text.find(nxt) if nxt else len(text)
        content = text[:cut].strip(strip_chars)
        if REQUIRED_LINES_AST[i]:
            content = content.split(" - ")[0].strip()   # trailing " - comment"
            if content.startswith("-"):                  # "- [ ... ]" list bullets
                content = content[1:].strip()
            try:
                content = ast.literal_eval(content)
            except SyntaxError as e:
                print(f"Encountered SyntaxError with content {content}: {e}")
                raise
        layout[line.rstrip(":")] = content
        # reconstructed response: for the last line the reference slices with the STALE position of the previous line
        # (`required_line_idx` is not updated when there is no next marker, utils/llm.py:228-251); reproduced as is
        keep = cut if nxt else prev_cut
        rebuilt += text[:keep + len(nxt)]
        text = text[keep + len(nxt):]
        prev_cut = cut
    layout["Prompt"] = prompt
    return layout, rebuilt


def get_parsed_layout_json_resp(prompt, llm_kwargs=None, max_partial_response_retries=1, override_response=None,
                                strip_chars=_STRIP, save_leading_text=True, **kwargs):
    """JSON template (utils/llm.py:277-328).  The reference parses with pyjson5; strict JSON is tried first and JSON5
    features LLMs actually emit (single quotes, trailing commas) through `ast.literal_eval`."""
    if override_response is None:
        raise NotImplementedError("lvd_b200 does not query an LLM: pass override_response (a cached response)")
    assert max_partial_response_retries == 1, "no partial queries are allowed in with JSON format templates"
    resp = override_response.strip(strip_chars)
    if resp.startswith("Response:"):
        resp = resp[len("Response:"):]
    resp = resp.strip(strip_chars)
    try:
        layout = json.loads(resp)
    except ValueError:
        try:
            layout = ast.literal_eval(resp.replace("true", "True").replace("false", "False").replace("null", "None"))
        except (ValueError, SyntaxError) as e:
            print(f"Encountered exception in parsing the response with content {resp}: {e}")
            raise ValueError(str(e)) from e
    layout["Prompt"] = prompt
    return layout, resp
