// HashTable.h -- drop-in for MetaGenomics/HashTable.h:20-36. The table itself lives in HBM
// (libogb K1); this class owns the GPU context and answers getListOfReads through ogb_hash_lookup.
#ifndef OGB_HOST_HASHTABLE_H_
#define OGB_HOST_HASHTABLE_H_

#include "Common.h"
#include "Dataset.h"

class HashTable
{
	private:
		Dataset *dataSet;
		UINT64 hashTableSize;						// slots in the device table
		UINT16 hashStringLength;					// minOverlap - 1 (HashTable.cpp:54)
		UINT64 numberOfHashCollision;				// kept for interface parity; the device table does not count
		ogb_context *context;
		bool ownsContext;
		map<string, vector<UINT64> *> lookupCache;	// lists handed out by getListOfReads stay valid
		friend class OverlapGraph;

	public:
		HashTable(void);							// GPU 0 (or $OGB_DEVICE)
		explicit HashTable(ogb_context *ctx);		// caller-owned context (multi-GPU ranks)
		~HashTable();
		bool insertDataset(Dataset *d, UINT64 minOverlapLength);
		vector<UINT64> *getListOfReads(string subString);
		UINT64 hashFunction(string subString);
		UINT64 getHashTableSize(void) { return hashTableSize; }
		UINT64 getHashStringLength() { return hashStringLength; }
		Dataset *getDataset(void) { return dataSet; }
		ogb_context *getContext(void) { return context; }
};

#endif
