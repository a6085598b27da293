/**
 * node_bench.ts -- times the UNMODIFIED reference (npm `bpe-tokenizer` 2.2.0, its `core.ts` class) on the workloads of
 * BASELINE.json, for users who have Node.  UNEXECUTED: Node, npm and tsc are absent from the authoring image and from the
 * GPU boxes (SURVEY.md section 8(c)), which is why bench.py --impl reference times the C++ restatement (oracle/int_oracle.cpp)
 * instead.  Run with:   npm i bpe-tokenizer@2.2.0 && npx ts-node baseline/node_bench.ts [bytes=16000000] [merges=8]
 *
 * The corpus is the same seeded Zipf text as bench.py's (bpe_tokenizer_b200/synth.py is the specification; this file
 * restates it with BigInt arithmetic): splitmix64 counter streams, 50 000 words of 2-10 letters (seed 42), Zipf weights
 * floor(2^40/(k+1)), documents '\r' + 8..64 words + '\n' (seed 43).  Output: one JSON line shaped like bench.py's.
 */
import { BPETokenizer } from 'bpe-tokenizer'

const MASK = (1n << 64n) - 1n
const GAMMA = 0x9e3779b97f4a7c15n
function mix(z: bigint): bigint {
  z = ((z ^ (z >> 30n)) * 0xbf58476d1ce4e5b9n) & MASK
  z = ((z ^ (z >> 27n)) * 0x94d049bb133111ebn) & MASK
  return z ^ (z >> 31n)
}
/** draw i of the stream seeded with `seed` */
const draw = (seed: bigint, i: number): bigint => mix((seed + BigInt(i + 1) * GAMMA) & MASK)

function wordList(vocab = 50000, seed = 42n): { words: string[]; cum: bigint[] } {
  const words: string[] = []
  let p = 0
  for (let k = 0; k < vocab; k++) {
    const len = 2 + Number(draw(seed, p++) % 9n)
    let w = ''
    for (let j = 0; j < len; j++) w += String.fromCharCode(97 + Number(draw(seed, p++) % 26n))
    words.push(w)
  }
  const cum: bigint[] = []
  let total = 0n
  for (let k = 0; k < vocab; k++) cum.push((total += (1n << 40n) / BigInt(k + 1)))
  return { words, cum }
}

function* documents(target_bytes: number, seed = 43n): Generator<string> {
  const { words, cum } = wordList()
  const total = cum[cum.length - 1]
  let bytes = 0
  for (let d = 0; bytes < target_bytes; d++) {
    const doc_seed = draw(seed, d)
    const n_words = 8 + Number(draw(doc_seed, 0) % 57n)
    const picked: string[] = []
    for (let i = 0; i < n_words; i++) {
      const r = draw(doc_seed, i + 1) % total
      let lo = 0
      let hi = cum.length // first k with cum[k] > r
      while (lo < hi) {
        const mid = (lo + hi) >> 1
        if (cum[mid] > r) hi = mid
        else lo = mid + 1
      }
      picked.push(words[lo])
    }
    const doc = '\r' + picked.join(' ') + '\n'
    bytes += doc.length
    yield doc
  }
}

const target_bytes = Number(process.argv[2] || 16_000_000)
const merges = Number(process.argv[3] || 8)
const tokenizer = new BPETokenizer()
let chars = 0
for (const doc of documents(target_bytes)) {
  tokenizer.addToCorpus(doc)
  chars += doc.length
}
const t0 = process.hrtime.bigint()
tokenizer.mergeUntil({ max_iterations: merges })
const seconds = Number(process.hrtime.bigint() - t0) / 1e9
const done = tokenizer.merge_tokens.length
// per-merge cost is linear in the corpus size: scale to the 1 GB workload the way bench.py's cpu_baseline does
console.log(
  JSON.stringify({
    impl: 'reference',
    metric: 'mergeUntil merges/sec',
    value: (done / seconds) * (chars / 1_000_000_287),
    unit: 'merges/s',
    cpu_baseline: {
      kind: 'reference',
      cores: 1,
      sample: `node ${process.version}: first ${done} merges on the first ${chars} chars took ${seconds.toFixed(2)} s, scaled by size`,
    },
    merges: tokenizer.merge_tokens.map(([a, b, c]) => [a.index, b.index, c.weight]),
  }),
)
