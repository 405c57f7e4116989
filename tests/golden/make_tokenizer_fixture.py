"""Builds tests/golden/llama3_style_tokenizer/: a REAL `tokenizers`-backed HF tokenizer (PreTrainedTokenizerFast) with a
Llama-3.1-style Jinja chat template, sized for the tiny configuration (512 ids; `preprocess` adds the 7 speech / latency
tokens -> 519 = the tiny model's vocabulary).

What it reproduces of the Llama-3.1-Instruct tokenizer (not in this image, model files are gated):
  * the special tokens the agent and `preprocess` look up: <|begin_of_text|>, <|start_header_id|>, <|end_header_id|>,
    <|eot_id|>, <|end_of_text|>, <|eom_id|>, <|finetune_right_pad_id|>, and the role words `system` / `user` / `assistant`
    as single tokens (model/llm.py:176-178 converts them to ids);
  * the chat template's layout: BOS, then for every message  <|start_header_id|> role <|end_header_id|> "\n\n" content
    <|eot_id|>; a conversation WITHOUT a system message still gets the default system turn
    "Cutting Knowledge Date: December 2023\nToday Date: 26 Jul 2024\n\n" - with BOS and the header that is exactly 25
    tokens before its <|eot_id|>, the prefix agents/infinisst.py:262-264 strips with `[:, 25:]`;
  * "\n\n" and "\n" as single tokens.
Run:  python tests/golden/make_tokenizer_fixture.py   (the directory it writes is committed)."""
import json
import os

from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers
from transformers import PreTrainedTokenizerFast

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "llama3_style_tokenizer")
VOCAB = 512

SPECIAL = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eom_id|>", "<|eot_id|>",
           "<|finetune_right_pad_id|>"]
# pieces as the Llama-3 BPE cuts the default header: words carry their leading space, digits come in groups of <= 3 after
# a lone space, "Cutting" is two pieces ("Cut", "ting") - 20 pieces for the two date lines, 25 with BOS and the header
WORDS = ["system", "user", "assistant", "\n\n", "\n", " ", ":", ".", ",", "Cut", "ting", " Knowledge", " Date", " December",
         "202", "3", "4", "Today", "26", " Jul", "Translate", " the", " following", " speech", " from", " to", " with",
         " latency", " English", " German", " Chinese", " Spanish", "(", ")", "（", "Hallo", " Welt", " und", " ist", " das",
         " ein", " der", " die"]

# Llama-3.1's template, reduced to the branches the agent exercises (no tools): BOS; the system turn (default date
# header, then the system message if there is one); every other message; no generation prompt (the agent appends an
# empty assistant message and drops the final <|eot_id|> itself, agents/infinisst.py:254-260)
TEMPLATE = (
    "{{- bos_token }}"
    "{%- if messages[0]['role'] == 'system' %}"
    "{%- set system_message = messages[0]['content'] | trim %}{%- set messages = messages[1:] %}"
    "{%- else %}{%- set system_message = '' %}{%- endif %}"
    "{{- '<|start_header_id|>system<|end_header_id|>\\n\\n' }}"
    "{{- 'Cutting Knowledge Date: December 2023\\n' }}{{- 'Today Date: 26 Jul 2024\\n\\n' }}"
    "{{- system_message }}{{- '<|eot_id|>' }}"
    "{%- for message in messages %}"
    "{{- '<|start_header_id|>' + message['role'] + '<|end_header_id|>\\n\\n' + message['content'] | trim + '<|eot_id|>' }}"
    "{%- endfor %}"
    "{%- if add_generation_prompt %}{{- '<|start_header_id|>assistant<|end_header_id|>\\n\\n' }}{%- endif %}"
)


def main():
    vocab = {}
    for w in WORDS:
        vocab[w] = len(vocab)
    i = 0
    n_plain = VOCAB - len(SPECIAL)
    while len(vocab) < n_plain - 1:
        vocab[f"w{i}"] = len(vocab)
        i += 1
    vocab["<unk>"] = len(vocab)
    assert len(vocab) == n_plain
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    # "\n\n" before "\n", words with their leading space, lone spaces, digit groups, single punctuation marks
    tok.pre_tokenizer = pre_tokenizers.Split(Regex(r"\n\n|\n|Cut(?=ting)| ?[A-Za-z]+| |[0-9]{1,3}|[^\sA-Za-z0-9]"), behavior="isolated")
    tok.decoder = decoders.Fuse()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<|begin_of_text|>", eos_token="<|eot_id|>",
                                   unk_token="<unk>", padding_side="right", clean_up_tokenization_spaces=False)
    fast.add_special_tokens({"additional_special_tokens": [t for t in SPECIAL if t not in ("<|begin_of_text|>", "<|eot_id|>")]})
    assert len(fast) == VOCAB, len(fast)
    fast.chat_template = TEMPLATE
    os.makedirs(OUT, exist_ok=True)
    fast.save_pretrained(OUT)
    # what the agent relies on (it sets the pad token itself after loading, agents/infinisst.py:140)
    fast = PreTrainedTokenizerFast.from_pretrained(OUT)
    fast.pad_token = "<|finetune_right_pad_id|>"
    ids = fast.apply_chat_template([[{"role": "user", "content": "x"}]], return_tensors="pt", padding=True,
                                   truncation=False, add_special_tokens=False)
    ids = ids["input_ids"] if hasattr(ids, "keys") else ids
    row = ids[0].tolist()
    eot = fast.convert_tokens_to_ids("<|eot_id|>")
    assert row.index(eot) == 25, (row.index(eot), fast.convert_ids_to_tokens(row))
    print("default system header:", fast.convert_ids_to_tokens(row[:26]))
    print("wrote", OUT, sorted(os.listdir(OUT)), "vocab", len(fast))


if __name__ == "__main__":
    main()
