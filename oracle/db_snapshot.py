"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the snapshot import/export of the reference's sqlite twin
(``BPETokenizerDB.fromJSON`` db/core.ts:165-201, ``toJSON`` db/core.ts:150-162, schema db/migration.ts:3-30), over
Python's stdlib sqlite3.  The sqlite tokenizer itself is out of scope; this file exists
to prove that the snapshots this repo's ``toJSON`` emits import into the reference's database unchanged
(pinned by the reference's own test db/core.spec.ts:32-41)."""
from __future__ import annotations

import sqlite3

SCHEMA = """
create table token (id integer primary key, chars text not null, weight integer not null, original_weight integer not null, code text not null);
create table char_token (id integer primary key);
create table merge (id integer primary key, a_id integer not null references token(id), b_id integer not null references token(id),
                    c_id integer not null references token(id));
"""  # db/migration.ts:11-30 (timestamps omitted: not part of the snapshot)


class DBSnapshot:
    def __init__(self):
        self.db = sqlite3.connect(":memory:")
        self.db.executescript(SCHEMA)

    def from_json(self, json: dict) -> None:  # db/core.ts:165-201
        if json.get("version") != 2 or not isinstance(json.get("token_table"), list) or not isinstance(json.get("merge_codes"), list):
            raise ValueError("invalid format")  # db/core.ts:166-171
        char_count = json["char_count"]
        self.db.executescript("delete from merge; delete from char_token; delete from token;")  # reset(), db/core.ts:141-146
        code_to_token = {}
        token_id = 0
        for chars, weight, original_weight in json["token_table"]:
            token_id += 1  # 1-based ids; code = fromCodePoint(token_id), db/core.ts:177-179 == core.ts:149 (index + 1)
            code = chr(token_id)
            self.db.execute("insert into token (id, chars, weight, original_weight, code) values (?,?,?,?,?)",
                            (token_id, chars, weight, original_weight, code))
            if token_id <= char_count:
                self.db.execute("insert into char_token (id) values (?)", (token_id,))  # db/core.ts:188-190
            code_to_token[code] = token_id
        for a_code, b_code, c_code in json["merge_codes"]:
            # db/core.ts:193-198: a missing code raises (TypeError in the reference)
            self.db.execute("insert into merge (a_id, b_id, c_id) values (?,?,?)", (code_to_token[a_code], code_to_token[b_code], code_to_token[c_code]))
        self.db.commit()

    def to_json(self) -> dict:  # db/core.ts:150-162 with the selects of :87-103
        token_table = [list(r) for r in self.db.execute("select chars, weight, original_weight from token order by token.id asc")]
        merge_codes = [list(r) for r in self.db.execute(
            "select a.code, b.code, c.code from merge inner join token a on a.id = merge.a_id inner join token b on b.id = merge.b_id "
            "inner join token c on c.id = merge.c_id order by merge.id asc")]
        char_count = self.db.execute("select count(*) from char_token").fetchone()[0]
        return {"version": 2, "char_count": char_count, "token_table": token_table, "merge_codes": merge_codes}
