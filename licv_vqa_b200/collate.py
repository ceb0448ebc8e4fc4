"""The collator's output contract (reference icv_src/icv_datamodule.py:73-130), for pre-tokenised
samples.

The reference tokenises four prompt variants per batch through lmm_icl_interface's processor
(query with label, query without label, in-context examples alone, examples + query - the
32-shot prompt is tokenised and image-preprocessed TWICE, :89-103) and derives two lengths from
pad/BOS counts.  The hot path only consumes the resulting dict:

    query_inputs      {"input_ids" [B,Tq], "attention_mask" [B,Tq]}     student prompt (+ EOS)
    inputs            {"input_ids" [B,Tt], "attention_mask" [B,Tt]}     teacher prompt = ICE ++ query
    in_context_length [B] = #non-pad(ICE) + #non-pad-non-BOS(query_x)    (:104-121)
    query_x_length    [B] = #non-pad(query_x)                            (:123-126)

`collate_token_ids` builds exactly that from token-id lists (any tokenizer; the examples are
tokenised once), `check_batch_contract` verifies the invariant the loss relies on - both masks of
`VQAICVModule.get_mask` select the same number of rows (icv_module.py:84-85,108-111) - on the host.
Integer work only; no device code.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


def _pad(rows: Sequence[Sequence[int]], pad_id: int, side: str):
    T = max(len(r) for r in rows)
    ids = torch.full((len(rows), T), pad_id, dtype=torch.long)
    att = torch.zeros((len(rows), T), dtype=torch.long)
    for b, r in enumerate(rows):
        n = len(r)
        if n == 0:
            continue
        if side == "right":
            ids[b, :n] = torch.tensor(r, dtype=torch.long)
            att[b, :n] = 1
        else:
            ids[b, T - n:] = torch.tensor(r, dtype=torch.long)
            att[b, T - n:] = 1
    return ids, att


def collate_token_ids(query_ids: List[List[int]], query_x_ids: List[List[int]],
                      ice_ids: List[List[int]], pad_token_id: int, bos_token_id: Optional[int],
                      eos_token_id: Optional[int] = None, padding_side: str = "right",
                      input_ids_field: str = "input_ids", max_length: Optional[int] = None) -> Dict:
    """query_ids: BOS + question + answer; query_x_ids: BOS + question (no answer); ice_ids: BOS +
    the k in-context examples.  The teacher prompt is what tokenising `ice + query` gives: the
    examples followed by the query WITHOUT a second BOS.  EOS is appended to the two prompts that
    carry the answer (`add_eos_token=True`, icv_datamodule.py:80-85,96-103) when `eos_token_id` is
    given; `max_length` truncates like `truncation=True`."""
    if not (len(query_ids) == len(query_x_ids) == len(ice_ids)):
        raise ValueError("query_ids, query_x_ids and ice_ids must have one entry per sample")
    if padding_side not in ("right", "left"):
        raise ValueError("padding_side must be 'right' or 'left'")
    tail = [eos_token_id] if eos_token_id is not None else []

    def no_bos(r):
        return list(r[1:]) if (bos_token_id is not None and len(r) and r[0] == bos_token_id) else list(r)

    student = [list(q) + tail for q in query_ids]
    teacher = [list(i) + no_bos(q) + tail for i, q in zip(ice_ids, query_ids)]
    if max_length is not None:
        student = [r[:max_length] for r in student]
        teacher = [r[:max_length] for r in teacher]
    q_ids, q_att = _pad(student, pad_token_id, padding_side)
    t_ids, t_att = _pad(teacher, pad_token_id, padding_side)

    def count(rows, skip_bos):
        return torch.tensor([sum(1 for x in r if x != pad_token_id and not (skip_bos and x == bos_token_id))
                             for r in rows], dtype=torch.long)

    in_context_length = count(ice_ids, False) + count(query_x_ids, True)
    query_x_length = count(query_x_ids, False)
    return {
        "query_inputs": {input_ids_field: q_ids, "attention_mask": q_att},
        "inputs": {input_ids_field: t_ids, "attention_mask": t_att},
        "in_context_length": in_context_length,
        "query_x_length": query_x_length,
    }


def check_batch_contract(batch: Dict, pad_token_id: int, input_ids_field: str = "input_ids") -> int:
    """Raise ValueError unless the batch has the collator's four keys and both get_mask calls
    select the same rows count per batch (what `stu - tea` at icv_module.py:126-131 needs);
    returns N, the number of KL rows."""
    for key in ("query_inputs", "inputs", "in_context_length", "query_x_length"):
        if key not in batch:
            raise ValueError(f"batch is missing '{key}' (icv_datamodule.py:125-130)")
    q = batch["query_inputs"][input_ids_field]
    t = batch["inputs"][input_ids_field]
    if q.shape[0] != t.shape[0] or q.shape[0] != batch["query_x_length"].shape[0]:
        raise ValueError("student and teacher batches differ")
    pos_q = torch.arange(q.shape[1])[None].to(q.device)
    pos_t = torch.arange(t.shape[1])[None].to(t.device)
    n_s = int(((pos_q >= batch["query_x_length"].to(q.device)[:, None]) & (q != pad_token_id)).sum())
    n_t = int(((pos_t >= batch["in_context_length"].to(t.device)[:, None]) & (t != pad_token_id)).sum())
    if n_s != n_t:
        raise ValueError(f"student mask selects {n_s} rows, teacher mask {n_t}: the two prompts "
                         "do not end in the same answer tokens")
    return n_s
