// `@noble/bls12-381/math` (package.json:46-51 of the reference): the value classes and constants are not on the device hot
// path; they are re-exported unchanged so that `import { Fp12 } from '.../math'` keeps resolving.
export * from '@noble/bls12-381/math';
